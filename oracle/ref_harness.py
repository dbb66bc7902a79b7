"""Drive the REFERENCE's own Python (read-only, /root/reference) as the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only usable where the reference tree exists (the build container): tools/make_golden.py runs it in a fresh process and
commits the resulting input/output vectors under tests/golden/.  Nothing under tests/ marked gpu, smoke() or bench.py
imports this module.

How: `envs.t1.T1` is instantiated with `T1.__new__` (its constructor needs Isaac Gym), the Gym tensor API is replaced by
a stub that hands back caller-owned tensors (oracle/shims/isaacgym), `_init_buffers()` and `_prepare_reward_function()`
of the reference then build every view exactly as in training, state is overwritten from the fixture, and the
reference's real `step()` / `reset()` run with "physics = identity" (gym.simulate is a no-op).  Random draws
(torch.randn_like / rand_like / rand / randint / randperm) are served from an injection table with the same slot
layout the CUDA kernels read (include/b200_t1.h: b200_t1_inject_rng), so all three implementations consume identical
samples.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
N_SLOTS = 21


def reference_dir():
    for cand in (os.environ.get("B200_REF_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "envs", "t1.py")):
            return cand
    return None


def import_reference():
    """returns (t1_module, terrain_module, utils_module, model_module) of the reference; mutates sys.path/sys.modules"""
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference tree not found (set B200_REF_DIR)")
    for name in list(sys.modules):
        if name in ("envs", "utils", "isaacgym") or name.startswith(("envs.", "utils.", "isaacgym.")):
            del sys.modules[name]
    sys.path[:0] = [os.path.join(HERE, "shims"), ref]
    import envs.t1 as t1_mod
    import utils.model as model_mod
    import utils.terrain as terrain_mod
    import utils.utils as utils_mod

    assert os.path.abspath(t1_mod.__file__).startswith(os.path.abspath(ref)), t1_mod.__file__
    return t1_mod, terrain_mod, utils_mod, model_mod


class NullGym:
    """Gym API stub: acquire_* return the tensors the harness owns, everything else is a no-op."""

    def __init__(self, tensors):
        self._t = tensors

    def acquire_actor_root_state_tensor(self, sim):
        return self._t["root_states"]

    def acquire_dof_state_tensor(self, sim):
        return self._t["dof_state"]

    def acquire_net_contact_force_tensor(self, sim):
        return self._t["contact_forces"]

    def acquire_rigid_body_state_tensor(self, sim):
        return self._t["body_states"]

    def __getattr__(self, name):
        return lambda *a, **k: None


# ---- injection table helpers (slot layout of include/b200_t1.h) -----------------------------------------------------
def make_table(n, seed, still_envs=None, still_proportion=0.1):
    """random table uint32 [N_SLOTS, 12, n]: words, uniforms in [0,1), standard normals.  The reset-dof slots (0-2) hold
    the same values for every env (reference quirk: one [1,12] draw per call).  If still_envs (bool [n]) is given, the
    'still' uniform (slot 7, lane 0) is made < still_proportion exactly for those envs."""
    rng = np.random.default_rng(seed)
    words = rng.integers(0, 2 ** 32, size=(N_SLOTS, 4, n), dtype=np.uint64).astype(np.uint32)
    uni = rng.random((N_SLOTS, 4, n), dtype=np.float32)
    uni = np.minimum(uni, np.float32(1.0 - 2 ** -24))
    nrm = rng.standard_normal((N_SLOTS, 4, n)).astype(np.float32)
    for s in range(3):
        words[s] = words[s][:, :1]
        uni[s] = uni[s][:, :1]
        nrm[s] = nrm[s][:, :1]
    if still_envs is not None:
        u = uni[7, 0]
        u[still_envs] = np.float32(still_proportion) * np.float32(0.5) * u[still_envs]
        u[~still_envs] = np.float32(still_proportion) + (np.float32(1.0) - np.float32(still_proportion)) * u[~still_envs] * np.float32(0.999)
        uni[7, 0] = u
    table = np.zeros((N_SLOTS, 12, n), dtype=np.uint32)
    table[:, 0:4] = words
    table[:, 4:8] = uni.view(np.uint32)
    table[:, 8:12] = nrm.view(np.uint32)
    return table


def tbl_u(table, slot, lane):
    return table[slot, 4 + lane].view(np.float32)


def tbl_n(table, slot, lane):
    return table[slot, 8 + lane].view(np.float32)


def tbl_w(table, slot, lane):
    return table[slot, lane]


class Injector:
    """serves the reference's torch RNG calls from the table, keyed by (phase, function, call index)"""

    def __init__(self, table, n):
        self.table, self.n = table, n
        self.phase, self.count, self.env_ids = None, {}, None
        self.log = []

    def enter(self, phase, env_ids=None):
        self.phase, self.count, self.env_ids = phase, {}, env_ids

    def _k(self, fn):
        k = self.count.get(fn, 0)
        self.count[fn] = k + 1
        self.log.append((self.phase, fn, k))
        return k

    def _cols(self, getter, slot0, lanes, ids=None, per_slot=4):
        """stack lanes (flat index i -> slot slot0 + i // 4, lane i % 4) -> float32 [len(ids) or n, lanes]"""
        cols = [getter(self.table, slot0 + i // per_slot, i % per_slot) for i in range(lanes)]
        a = np.stack(cols, axis=1)
        if ids is not None:
            a = a[ids]
        return torch.from_numpy(np.ascontiguousarray(a))

    def randn_like(self, t, **kw):
        k = self._k("randn_like")
        ph, ids = self.phase, self.env_ids
        if ph == "kick":
            return self._cols(tbl_n, 8 + k, 3).reshape(t.shape)
        if ph == "push":
            return self._cols(tbl_n, 10 + k, 3).reshape(t.shape)
        if ph == "reset":
            if k == 0:  # init_dof_pos: shaped like default_dof_pos [1,12] -> shared by every env of the call
                return self._cols(tbl_n, 0, 12, ids=ids[:1]).reshape(t.shape)
            if k == 1:  # init_base_lin_vel_xy [len,2]
                return self._cols(tbl_n, 4, 2, ids=ids).reshape(t.shape)
        if ph == "obs":
            start, width = [(0, 3), (3, 3), (6, 12), (18, 12), (30, 3), (33, 1)][k]
            cols = [tbl_n(self.table, 12 + (start + i) // 4, (start + i) % 4) for i in range(width)]
            return torch.from_numpy(np.ascontiguousarray(np.stack(cols, axis=1))).reshape(t.shape)
        raise AssertionError(f"unexpected randn_like in phase {ph} (call {k}, shape {tuple(t.shape)})")

    def rand_like(self, t, **kw):
        k = self._k("rand_like")
        if self.phase == "reset" and k == 0:  # init_base_pos_xy [len,2]
            return self._cols(tbl_u, 3, 2, ids=self.env_ids).reshape(t.shape)
        raise AssertionError(f"unexpected rand_like in phase {self.phase}")

    def rand(self, *shape, **kw):
        k = self._k("rand")
        ids = self.env_ids
        if self.phase == "reset" and k == 0:  # yaw
            return self._cols(tbl_u, 3, 3, ids=ids)[:, 2].reshape(*shape)
        if self.phase == "command":  # vx, vy, yaw rate, gait frequency: torch_rand_float(..., (len, 1))
            return self._cols(tbl_u, 6, 4, ids=ids)[:, k].reshape(*shape)
        raise AssertionError(f"unexpected rand in phase {self.phase}")

    def randint(self, low, high, size, **kw):
        self._k("randint")
        ids = self.env_ids
        if self.phase == "reset":
            w = tbl_w(self.table, 5, 0)[ids].astype(np.int64)
        elif self.phase == "command":
            w = tbl_w(self.table, 7, 1)[ids].astype(np.int64)
        else:
            raise AssertionError(f"unexpected randint in phase {self.phase}")
        return torch.from_numpy(low + w % (high - low)).reshape(size)

    def multinomial(self, probs, num_samples, replacement=False, **kw):
        """torch.multinomial cannot be fed samples: served as the inverse CDF of the same distribution on the table's uniform
        (slot 7, lane 2) - sequential fp32 running sum, first cell whose sum exceeds u * total (oracle/env_oracle.py, t1_env.cuh)"""
        self._k("multinomial")
        assert self.phase == "command" and replacement and num_samples == len(self.env_ids)
        p = probs.detach().numpy().astype(np.float32)
        cdf = np.add.accumulate(p, dtype=np.float32)
        target = (tbl_u(self.table, 7, 2)[self.env_ids].astype(np.float32) * cdf[-1]).astype(np.float32)
        return torch.from_numpy(np.minimum(np.searchsorted(cdf, target, side="right"), len(cdf) - 1).astype(np.int64))

    def randperm(self, m, **kw):
        self._k("randperm")
        assert self.phase == "command"
        ids = self.env_ids
        u = tbl_u(self.table, 7, 0)[ids]
        still = np.nonzero(u < np.float32(self.still_proportion))[0]
        rest = np.nonzero(~(u < np.float32(self.still_proportion)))[0]
        k = int(self.still_proportion * m)
        assert len(still) == k, f"table must mark exactly int(still_proportion * {m}) = {k} still envs, has {len(still)}"
        return torch.from_numpy(np.concatenate([still, rest]).astype(np.int64))


class patched_rng:
    def __init__(self, inj):
        self.inj = inj

    def __enter__(self):
        self.saved = {k: getattr(torch, k) for k in ("randn_like", "rand_like", "rand", "randint", "randperm", "multinomial")}
        for k in self.saved:
            setattr(torch, k, getattr(self.inj, k))

    def __exit__(self, *a):
        for k, v in self.saved.items():
            setattr(torch, k, v)


# ---- building a reference env from a fixture ---------------------------------------------------------------------------
STATE_KEYS_F = ["root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel", "last_dof_targets",
                "torques", "commands", "gait_frequency", "gait_process", "filtered_lin_vel", "filtered_ang_vel", "pushing_forces",
                "pushing_torques", "feet_pos", "feet_quat", "last_feet_pos", "dof_stiffness", "dof_damping", "dof_friction",
                "base_mass_scaled", "env_origins"]
STATE_KEYS_I = ["episode_length_buf", "cmd_resample_time", "delay_steps"]


def build_reference_env(mods, cfg, state, hf=None):
    """state: dict of numpy arrays (STATE_KEYS_*; [N, ...] row-major like the reference tensors)"""
    t1_mod, terrain_mod, _, _ = mods
    js = json.load(open(os.path.join(ROOT, "booster_gym_b200", "assets", "t1_model.json")))
    n = state["root_states"].shape[0]
    T = lambda a, dt=torch.float: torch.tensor(np.asarray(a), dtype=dt)  # noqa: E731
    body_states = torch.zeros(n, 13, 13)
    body_states[:, [6, 12], 0:3] = T(state["feet_pos"]).reshape(n, 2, 3)
    body_states[:, [6, 12], 3:7] = T(state["feet_quat"]).reshape(n, 2, 4)
    dof_state = torch.stack([T(state["dof_pos"]), T(state["dof_vel"])], dim=-1).reshape(n * 12, 2).contiguous()
    tensors = dict(root_states=T(state["root_states"]).clone(), dof_state=dof_state,
                   contact_forces=(T(state["contact_forces"]).reshape(n * 13, 3).clone() if "contact_forces" in state
                                   else torch.zeros(n * 13, 3)),
                   body_states=body_states.reshape(n * 13, 13))
    e = t1_mod.T1.__new__(t1_mod.T1)
    e.cfg, e.device, e.gym, e.sim = cfg, "cpu", NullGym(tensors), None
    e.viewer, e.camera, e.headless, e.up_axis_idx, e.enable_viewer_sync = None, None, True, 2, True
    terr = terrain_mod.Terrain.__new__(terrain_mod.Terrain)
    terr.terrain_cfg, terr.device, terr.type = cfg["terrain"], "cpu", cfg["terrain"]["type"]
    if terr.type == "trimesh":
        tc = cfg["terrain"]
        terr.env_width = tc["num_terrains"] * tc["terrain_width"]
        terr.env_length = tc["terrain_length"]
        terr.border_size = tc["border_size"]
        terr.horizontal_scale, terr.vertical_scale = tc["horizontal_scale"], tc["vertical_scale"]
        terr.border_pixels = int(terr.border_size / terr.horizontal_scale)
        terr.height_field_raw = hf
    e.terrain = terr
    e.num_envs, e.num_dofs, e.num_bodies = n, 12, 13
    e.dof_names = js["dof_names"]
    e.dof_pos_limits = torch.tensor(np.stack([js["urdf_lower"], js["urdf_upper"]], axis=1), dtype=torch.float)
    e.dof_vel_limits = torch.tensor(js["urdf_velocity"], dtype=torch.float)
    e.torque_limits = torch.tensor(js["urdf_effort"], dtype=torch.float)
    e.dof_stiffness, e.dof_damping, e.dof_friction = T(state["dof_stiffness"]), T(state["dof_damping"]), T(state["dof_friction"])
    names = js["body_names"]
    pen = []
    for key in cfg["rewards"]["penalize_contacts_on"]:
        pen.extend([s for s in names if key in s])
    e.penalized_contact_indices = torch.tensor([names.index(s) for s in pen], dtype=torch.long)
    term = []   # envs/t1.py:90-92
    for key in cfg["rewards"]["terminate_contacts_on"]:
        term.extend([s for s in names if key in s])
    e.termination_contact_indices = torch.tensor([names.index(s) for s in term], dtype=torch.long)
    e.base_indice = 0
    e.feet_indices = torch.tensor([6, 12], dtype=torch.long)
    st = cfg["init_state"]
    e.base_init_state = torch.tensor(st["pos"] + st["rot"] + st["lin_vel"] + st["ang_vel"], dtype=torch.float)
    e.env_origins = T(state["env_origins"])
    e.base_mass_scaled = T(state["base_mass_scaled"])
    e._init_buffers()
    e._prepare_reward_function()
    # overwrite what _init_buffers zero-initialised
    for key in ("actions", "last_actions", "last_dof_vel", "last_root_vel", "last_dof_targets", "torques", "commands",
                "gait_frequency", "gait_process", "filtered_lin_vel", "filtered_ang_vel"):
        getattr(e, key)[:] = T(state[key])
    e.pushing_forces[:, 0, :] = T(state["pushing_forces"])
    e.pushing_torques[:, 0, :] = T(state["pushing_torques"])
    e.last_feet_pos[:] = T(state["last_feet_pos"]).reshape(n, 2, 3)
    for key in STATE_KEYS_I:
        getattr(e, key)[:] = torch.tensor(state[key], dtype=torch.long)
    if "curriculum_prob" in state:       # command curriculum (envs/t1.py:255-262)
        e.curriculum_prob[:] = T(state["curriculum_prob"])
        e.env_curriculum_level[:] = torch.tensor(state["env_curriculum_level"], dtype=torch.long)
    return e


def instrument(e, inj, cfg):
    """wrap the reference methods so the injector knows which draw is being served; capture post-loop torques"""
    inj.still_proportion = cfg["commands"]["still_proportion"]
    captured = {}

    def wrap(name, phase, ids_fn=None):
        orig = getattr(e, name)

        def f(*a, **k):
            ids = ids_fn(*a, **k) if ids_fn else None
            inj.enter(phase, None if ids is None else ids.numpy())
            return orig(*a, **k)

        setattr(e, name, f)

    wrap("_kick_robots", "kick")
    wrap("_push_robots", "push")
    wrap("_reset_idx", "reset", lambda env_ids: env_ids)
    wrap("_resample_commands", "command", lambda: (e.episode_length_buf == e.cmd_resample_time).nonzero(as_tuple=False).flatten())
    wrap("_compute_observations", "obs")
    orig_render = e.render

    def render():
        captured["torques"] = e.torques.clone().numpy()
        captured["last_dof_targets"] = e.last_dof_targets.clone().numpy()
        captured["actions"] = e.actions.clone().numpy()
        return orig_render()

    e.render = render
    return captured


def snapshot(e):
    """every observable of the env after a call, as numpy"""
    n = e.num_envs
    out = dict(
        obs=e.obs_buf, priv=e.privileged_obs_buf, rew=e.rew_buf, reset_buf=e.reset_buf, time_out_buf=e.time_out_buf,
        extras_time_outs=e.extras.get("time_outs", torch.zeros(n, dtype=torch.bool)),
        root_states=e.root_states, dof_pos=e.dof_pos, dof_vel=e.dof_vel, last_actions=e.last_actions, last_dof_vel=e.last_dof_vel,
        last_root_vel=e.last_root_vel, last_dof_targets=e.last_dof_targets, commands=e.commands, gait_frequency=e.gait_frequency,
        gait_process=e.gait_process, base_lin_vel=e.base_lin_vel, base_ang_vel=e.base_ang_vel, projected_gravity=e.projected_gravity,
        filtered_lin_vel=e.filtered_lin_vel, filtered_ang_vel=e.filtered_ang_vel, pushing_forces=e.pushing_forces[:, 0, :],
        pushing_torques=e.pushing_torques[:, 0, :], feet_roll=e.feet_roll, feet_yaw=e.feet_yaw, feet_contact=e.feet_contact,
        last_feet_pos=e.last_feet_pos.reshape(n, 6), feet_pos=e.feet_pos.reshape(n, 6), episode_length_buf=e.episode_length_buf,
        cmd_resample_time=e.cmd_resample_time, delay_steps=e.delay_steps,
    )
    if e.cfg["commands"].get("curriculum"):
        out.update(curriculum_prob=e.curriculum_prob, env_curriculum_level=e.env_curriculum_level,
                   mean_lin_vel_level=torch.as_tensor(e.mean_lin_vel_level), mean_ang_vel_level=torch.as_tensor(e.mean_ang_vel_level),
                   max_lin_vel_level=torch.as_tensor(e.max_lin_vel_level), max_ang_vel_level=torch.as_tensor(e.max_ang_vel_level))
    res = {k: v.detach().clone().numpy() for k, v in out.items()}
    for name, v in e.extras["rew_terms"].items():
        res["term_" + name] = v.detach().clone().numpy()
    return res
