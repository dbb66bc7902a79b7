"""envs/T1.yaml dictionary -> the plain-C config structs of include/b200_t1.h.

Everything the reference derives from the YAML at construction time (envs/t1.py:33-137,187-292) is derived here once
on the host with the same Python arithmetic (Python floats are fp64; the result is rounded to fp32 exactly where the
reference's torch ops would round), so the kernels only see numbers.
"""
import math

import numpy as np

from . import _abi, robot

# body / DoF names after collapse_fixed_joints; also in assets/t1_model.json
_MODEL = None


def model_json():
    global _MODEL
    if _MODEL is None:
        _MODEL = robot.load_json()
    return _MODEL


def _rand(params):
    """apply_randomization() descriptor (utils/utils.py:5-30). Raises the reference's ValueErrors."""
    r = _abi.Rand()
    if params is None:
        r.enabled = 0
        return r
    dist, op = params["distribution"], params["operation"]
    if dist == "gaussian":
        r.dist = 0
        r.a, r.b = float(params["range"][0]), float(params["range"][1])  # mu + "var" * randn
    elif dist == "uniform":
        r.dist = 1
        lo, hi = float(params["range"][0]), float(params["range"][1])
        r.a, r.b = lo, hi - lo  # lower + (upper - lower) * rand, (upper - lower) formed in fp64 like Python
    else:
        raise ValueError(f"Invalid randomization distribution: {dist}")
    if op == "additive":
        r.op = 0
    elif op == "scaling":
        r.op = 1
    else:
        raise ValueError(f"Invalid randomization operation: {op}")
    r.enabled = 1
    return r


def _by_substring(table, names, default_key=None, what="value"):
    out = []
    for n in names:
        val, found = None, False
        for key in table.keys():
            if key == default_key:
                continue
            if key in n:  # later matches override earlier ones, like the reference loop (envs/t1.py:74-78)
                val, found = table[key], True
        if not found:
            if default_key is None:
                raise ValueError(f"PD gain of joint {n} were not defined")
            val = table[default_key]
        out.append(float(val))
    return out


def reward_terms(cfg):
    """[(name, scale*dt)] of the non-zero scales in YAML order (envs/t1.py:279-292)."""
    dt = cfg["control"]["decimation"] * cfg["sim"]["dt"]
    out = []
    for name, scale in cfg["rewards"]["scales"].items():
        if scale == 0:
            continue
        if name not in _abi.REW_NAMES:
            raise AttributeError(f"'T1' object has no attribute '_reward_{name}'")
        out.append((name, scale * dt))
    return out


def t1_config(cfg):
    """Fill a B200T1Config from the YAML dict."""
    js = model_json()
    dof_names = js["dof_names"]
    body_names = js["body_names"]
    c = _abi.T1Config()
    dt = cfg["control"]["decimation"] * cfg["sim"]["dt"]
    c.env_dt = dt
    c.decimation = int(cfg["control"]["decimation"])
    c.action_scale = cfg["control"]["action_scale"]
    c.clip_actions = cfg["normalization"]["clip_actions"]
    dq = _by_substring(cfg["init_state"]["default_joint_angles"], dof_names, default_key="default")
    kp = _by_substring(cfg["control"]["stiffness"], dof_names)
    kd = _by_substring(cfg["control"]["damping"], dof_names)
    for j in range(_abi.NU):
        c.default_dof_pos[j] = dq[j]
        c.kp_nominal[j] = kp[j]
        c.kd_nominal[j] = kd[j]
    st = cfg["init_state"]
    for k, v in enumerate(list(st["pos"]) + list(st["rot"]) + list(st["lin_vel"]) + list(st["ang_vel"])):
        c.init_root[k] = float(v)
    c.env_spacing = cfg["env"]["env_spacing"]
    nz = cfg["normalization"]
    c.norm_gravity, c.norm_lin_vel, c.norm_ang_vel = nz["gravity"], nz["lin_vel"], nz["ang_vel"]
    c.norm_dof_pos, c.norm_dof_vel, c.filter_weight = nz["dof_pos"], nz["dof_vel"], nz["filter_weight"]
    c.norm_push_force, c.norm_push_torque = nz["push_force"], nz["push_torque"]
    noise = cfg.get("noise", {}) or {}
    for key in ("gravity", "lin_vel", "ang_vel", "dof_pos", "dof_vel", "height"):
        setattr(c, "noise_" + key, _rand(noise.get(key)))
    rz = cfg["randomization"]
    for key in ("init_dof_pos", "init_base_pos_xy", "init_base_lin_vel_xy", "kick_lin_vel", "kick_ang_vel", "push_force",
                "push_torque", "dof_stiffness", "dof_damping", "dof_friction", "friction", "compliance", "restitution",
                "base_com", "base_mass", "other_com", "other_mass"):
        setattr(c, key, _rand(rz.get(key)))
    c.kick_interval = int(np.ceil(rz["kick_interval_s"] / dt))
    c.push_interval = int(np.ceil(rz["push_interval_s"] / dt))
    c.push_duration = int(np.ceil(rz["push_duration_s"] / dt))
    c.push_all_substeps = 1 if rz.get("push_all_substeps") else 0
    cm = cfg["commands"]
    for key in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "gait_frequency"):
        getattr(c, key)[0] = float(cm[key][0])
        getattr(c, key)[1] = float(cm[key][1])
    c.still_proportion = cm["still_proportion"]
    c.resample_lo = int(cm["resampling_time_s"][0] / dt)
    c.resample_hi = int(cm["resampling_time_s"][1] / dt)
    c.curriculum = 1 if cm.get("curriculum") else 0
    # command curriculum (envs/T1.yaml:123-134, envs/t1.py:391-435); the grid is sized from the levels even when it is off
    c.cur_lin_levels, c.cur_ang_levels = int(cm.get("lin_vel_levels", 0)), int(cm.get("ang_vel_levels", 0))
    c.cur_update_rate = float(cm.get("update_rate", 0.0))
    c.cur_res_x, c.cur_res_y = float(cm.get("lin_vel_x_resolution", 0.0)), float(cm.get("lin_vel_y_resolution", 0.0))
    c.cur_res_ang = float(cm.get("ang_vel_resolution", 0.0))
    c.cur_success_len = float(np.ceil(cfg["rewards"]["episode_length_s"] / dt) * (1 - cm.get("episode_length_toler", 0.0)))
    c.cur_tol_x, c.cur_tol_y = float(cm.get("lin_vel_x_toler", 0.0)), float(cm.get("lin_vel_y_toler", 0.0))
    c.cur_tol_yaw = float(cm.get("ang_vel_yaw_toler", 0.0))
    rw = cfg["rewards"]
    terms = reward_terms(cfg)
    c.n_rew = len(terms)
    for k, (name, scale) in enumerate(terms):
        c.rew_id[k] = _abi.REW_NAMES.index(name)
        c.rew_scale[k] = scale
    c.max_episode_length = int(np.ceil(rw["episode_length_s"] / dt))
    c.terminate_height = rw["terminate_height"]
    c.terminate_vel = rw["terminate_vel"]
    c.only_positive_rewards = 1 if rw["only_positive_rewards"] else 0
    c.tracking_sigma = rw["tracking_sigma"]
    c.base_height_target = rw["base_height_target"]
    c.soft_dof_vel_limit = rw["soft_dof_vel_limit"]
    c.soft_torque_limit = rw["soft_torque_limit"]
    c.swing_period = rw["swing_period"]
    c.feet_distance_ref = rw["feet_distance_ref"]
    # envs/t1.py:665-670 in the reference's fp32 tensor arithmetic
    lo = np.asarray(js["urdf_lower"], dtype=np.float32)
    hi = np.asarray(js["urdf_upper"], dtype=np.float32)
    half = np.float32(0.5 * (1 - rw["soft_dof_pos_limit"]))
    soft_lo = lo + half * (hi - lo)
    soft_hi = hi - half * (hi - lo)
    for j in range(_abi.NU):
        c.dof_pos_soft_lower[j] = float(soft_lo[j])
        c.dof_pos_soft_upper[j] = float(soft_hi[j])
        c.dof_vel_limits[j] = js["urdf_velocity"][j]
        # envs/t1.py clips to the URDF effort limits Isaac Gym reports; play_mujoco.py:753-755 (BASELINE configs[0]) clips to the MJCF
        # actuator ctrlrange, which differs for hip roll (45 vs 30 N m) and knee (65 vs 60) - SURVEY 8a note 10.  asset.effort_limits: "mjcf"
        # selects the latter (tests/test_gpu_closed_loop.py); the shipped YAML has no such key = the training path's URDF limits.
        c.torque_limits[j] = (float(js["ctrlrange"][j][1]) if str(cfg.get("asset", {}).get("effort_limits", "urdf")).lower() == "mjcf"
                              else js["urdf_effort"][j])
    pen = 0
    for name in rw["penalize_contacts_on"]:
        for b, bn in enumerate(body_names):
            if name in bn:
                pen |= 1 << b
    term = 0
    for name in rw["terminate_contacts_on"]:
        for b, bn in enumerate(body_names):
            if name in bn:
                term |= 1 << b
    c.penalized_body_mask, c.termination_body_mask = pen, term
    tr = cfg["terrain"]
    if tr["type"] == "plane":
        c.terrain_type = 0
    elif tr["type"] == "trimesh":
        c.terrain_type = 1
    else:
        raise ValueError(f"Invalid terrain type: {tr['type']}")
    c.horizontal_scale = tr.get("horizontal_scale", 0.1)
    c.vertical_scale = tr.get("vertical_scale", 0.005)
    c.border_size = tr.get("border_size", 0.0)
    c.border_pixels = int(c.border_size / tr.get("horizontal_scale", 0.1)) if c.terrain_type == 1 else 0
    c.env_width = tr.get("num_terrains", 0) * tr.get("terrain_width", 0.0)
    c.env_length = tr.get("terrain_length", 0.0)
    c.terrain_friction = tr["static_friction"]
    return c


def ppo_config(cfg, num_envs, world_size=1, env_base=0):
    a = cfg["algorithm"]
    p = _abi.PpoConfig()
    p.gamma, p.lam = a["gamma"], a["lam"]
    p.e_clip = 0.2  # utils/utils.py:47 default argument
    p.bound_coef, p.entropy_coef, p.desired_kl = a["bound_coef"], a["entropy_coef"], a["desired_kl"]
    p.max_grad_norm = 1.0  # utils/runner.py:164
    p.lr_min, p.lr_max, p.lr_factor = 1e-5, 1e-2, 1.5  # utils/runner.py:175-178
    p.adam_beta1, p.adam_beta2, p.adam_eps = 0.9, 0.999, 1e-8  # torch.optim.Adam defaults (utils/runner.py:33)
    p.horizon = int(cfg["runner"]["horizon_length"])
    p.num_envs = int(num_envs)
    p.world_size = int(world_size)
    p.env_base = int(env_base)
    return p


def feet_edge_pos(cfg):
    return [list(map(float, row)) for row in cfg["asset"]["feet_edge_pos"]]


def sim_params(cfg):
    g = cfg["sim"]["gravity"]
    if cfg["sim"]["up_axis"] != "z":
        if cfg["sim"]["up_axis"] == "y":
            raise NotImplementedError("up_axis 'y' is not supported by the B200 T1 physics (the T1 model is z-up)")
        raise ValueError(f"Invalid physics up-axis: {cfg['sim']['up_axis']}")
    return dict(dt=float(cfg["sim"]["dt"]), gravity=-float(g[2]))


def env_origins(cfg, total_envs, terrain_heights=None):
    """_get_env_origins (envs/t1.py:169-185) for the GLOBAL env layout, float32 [total_envs, 3]."""
    n = int(total_envs)
    out = np.zeros((n, 3), dtype=np.float32)
    if cfg["terrain"]["type"] == "plane":
        num_cols = np.floor(np.sqrt(n))
        num_rows = np.ceil(n / num_cols)
        xx, yy = np.meshgrid(np.arange(int(num_rows)), np.arange(int(num_cols)), indexing="ij")
        spacing = cfg["env"]["env_spacing"]
        out[:, 0] = (np.float32(spacing) * xx.flatten()[:n].astype(np.float32))
        out[:, 1] = (np.float32(spacing) * yy.flatten()[:n].astype(np.float32))
    else:
        tr = cfg["terrain"]
        env_width = tr["num_terrains"] * tr["terrain_width"]
        env_length = tr["terrain_length"]
        num_cols = max(1.0, np.floor(np.sqrt(n * env_length / env_width)))
        num_rows = np.ceil(n / num_cols)
        xx, yy = np.meshgrid(np.arange(int(num_rows)), np.arange(int(num_cols)), indexing="ij")
        # torch: python_float / numpy_float -> fp64 scalar, times an int64 tensor (+1) -> fp32 result
        out[:, 0] = (np.float32(env_width / (num_rows + 1)) * (xx.flatten()[:n] + 1).astype(np.float32))
        out[:, 1] = (np.float32(env_length / (num_cols + 1)) * (yy.flatten()[:n] + 1).astype(np.float32))
        if terrain_heights is not None:
            out[:, 2] = terrain_heights(out)
    return out
