"""helpers shared by the golden-fixture tests"""
import copy
import json
import os

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

STATE_KEYS = ["root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel", "last_dof_targets",
              "torques", "commands", "gait_frequency", "gait_process", "filtered_lin_vel", "filtered_ang_vel", "pushing_forces",
              "pushing_torques", "feet_pos", "feet_quat", "last_feet_pos", "dof_stiffness", "dof_damping", "dof_friction",
              "base_mass_scaled", "env_origins", "episode_length_buf", "cmd_resample_time", "delay_steps"]

# outputs compared after a step: (key, kind)  kind: "exact" bit-exact, "float" 1e-5-relative, "angle" transcendental (1e-5 abs)
OUT_FLOAT = ["obs", "priv", "rew", "root_states", "dof_pos", "dof_vel", "last_actions", "last_dof_vel", "last_root_vel",
             "last_dof_targets", "commands", "gait_frequency", "gait_process", "base_lin_vel", "base_ang_vel", "projected_gravity",
             "filtered_lin_vel", "filtered_ang_vel", "pushing_forces", "pushing_torques", "feet_roll", "feet_yaw", "last_feet_pos",
             "feet_pos"]
OUT_EXACT = ["reset_buf", "time_out_buf", "extras_time_outs", "feet_contact", "episode_length_buf", "cmd_resample_time", "delay_steps"]


def load_cfg(terrain):
    cfg = yaml.load(open(os.path.join(ROOT, "envs", "T1.yaml")).read(), Loader=yaml.FullLoader)
    cfg = copy.deepcopy(cfg)
    cfg["terrain"]["type"] = terrain
    cfg["basic"]["headless"] = True
    return cfg


def fixture_cfg(name, terrain):
    """config a fixture was produced with (tools/make_golden.py)"""
    cfg = load_cfg(terrain)
    if name == "env_step_contacts.npz":   # contact termination is off in the shipped YAML (terminate_contacts_on: [])
        cfg["rewards"]["terminate_contacts_on"] = ["Trunk", "Shank"]
    return cfg


def model_json():
    return json.load(open(os.path.join(ROOT, "booster_gym_b200", "assets", "t1_model.json")))


def load(name):
    return np.load(os.path.join(GOLD, name))


CURRICULUM_KEYS = ["curriculum_prob", "env_curriculum_level"]


def step_inputs(z):
    """state dict for the post-physics half: the fixture's input state with what the reference's (identity-physics)
    decimation loop left behind: clipped actions, mean torques, last_dof_targets"""
    st = {k: z["in_" + k].copy() for k in STATE_KEYS}
    for k in CURRICULUM_KEYS:
        if "in_" + k in z.files:
            st[k] = z["in_" + k].copy()
    if "in_contact_forces" in z.files:   # the kernels consume |F_b| > 1 as one bit per body (state rows contact_mask)
        cf = z["in_contact_forces"]
        st["contact_forces"] = cf.copy()
        flags = np.sqrt(np.sum(cf * cf, axis=2, dtype=np.float32)) > np.float32(1.0)
        mask = (flags.astype(np.int64) << np.arange(13)[None, :]).sum(axis=1)
        left = mask & 0x7F            # trunk + left leg bits travel in the left lane's row, the right leg's in the other
        st["contact_mask"] = np.stack([left, mask & ~0x7F], axis=1)
    st["actions"] = z["post_loop_actions"].copy()
    st["torques"] = z["post_loop_torques"].copy()
    st["last_dof_targets"] = z["post_loop_last_dof_targets"].copy()
    return st


def close(a, b, rel=1e-5, atol=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), 1.0) * 1.0 + atol)
