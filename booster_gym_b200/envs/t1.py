"""T1 task: the Booster T1 environment API (mirror of envs/t1.py:24-603) driving the CUDA kernels of libb200t1.so.

Same surface as the reference class - `T1(cfg)`, `.reset() -> (obs, extras)`, `.step(actions) -> (obs, rew, done,
extras)` with `extras["privileged_obs"|"time_outs"|"rew_terms"]`, attributes `num_envs / num_obs / num_privileged_obs /
num_actions / dt / curriculum_prob / *_vel_level` and the per-env state tensors by their reference names - but every
tensor operation between the two calls is one of three kernels (physics, post-physics, time-out finalise) launched on
torch's current stream.  Per-env state is one structure-of-arrays float tensor [F_ROWS, N] and one int32 tensor
[I_ROWS, N]; the reference-named attributes are transposed VIEWS into it, so reads and in-place writes work as before.
"""
import ctypes as C

import numpy as np
import torch

from .. import _abi, _lib, config, robot
from .base_task import BaseTask


class T1(BaseTask):

    def __init__(self, cfg, env_index_base=0, total_envs=None, height_field=None):
        super().__init__(cfg)
        if height_field is not None:  # a given int16 heightfield instead of the generated one (BASELINE config 3: an input)
            if self.terrain.type == "plane":
                raise ValueError("height_field given but terrain.type is 'plane'")
            hf = np.ascontiguousarray(height_field, dtype=np.int16)
            if hf.shape != self.terrain.height_field_raw.shape:
                raise ValueError(f"height_field must have shape {self.terrain.height_field_raw.shape}")
            self.terrain.height_field_raw = hf
        self._lib = _lib.load()
        self._env_index_base = int(env_index_base)
        self._create_envs(total_envs)
        self._init_buffers()
        self._prepare_reward_function()

    # ---- construction ------------------------------------------------------------------------------------------
    def _create_envs(self, total_envs):
        cfg = self.cfg
        self.num_envs = int(cfg["env"]["num_envs"])
        self._total_envs = int(total_envs) if total_envs is not None else self.num_envs
        js = config.model_json()
        self.num_dofs = _abi.NU
        self.num_bodies = _abi.NB
        self.dof_names = list(js["dof_names"])
        self.body_names = list(js["body_names"])
        dev = self.device
        self.dof_pos_limits = torch.tensor(np.stack([js["urdf_lower"], js["urdf_upper"]], axis=1), dtype=torch.float, device=dev)
        self.dof_vel_limits = torch.tensor(js["urdf_velocity"], dtype=torch.float, device=dev)
        self.base_indice = self.body_names.index(cfg["asset"]["base_name"])
        self.feet_indices = torch.tensor([self.body_names.index(n) for n in cfg["asset"]["foot_names"]], dtype=torch.long, device=dev)
        if self.feet_indices.tolist() != js["feet_bodies"]:
            raise ValueError("asset.foot_names must name the two foot links of the T1 model")

        self._c_cfg = config.t1_config(cfg)
        self.torque_limits = torch.tensor(list(self._c_cfg.torque_limits), dtype=torch.float, device=dev)   # URDF efforts, or the MJCF ctrlrange (asset.effort_limits)
        sp = config.sim_params(cfg)
        # asset.self_collisions is Isaac Gym's collision FILTER: 0 = leg-leg contacts enabled, 1 = disabled (envs/T1.yaml:69)
        self._c_model = robot.model_f(foot_corner=config.feet_edge_pos(cfg), dt=sp["dt"], gravity=sp["gravity"],
                                      terrain_friction=float(cfg["terrain"]["static_friction"]),
                                      self_contact=int(cfg["asset"].get("self_collisions", 0)) == 0)
        hf = self.terrain.height_field_raw
        handle = C.c_void_p()
        if hf is not None:
            hf = np.ascontiguousarray(hf, dtype=np.int16)
            hf_ptr, rows, cols = hf.ctypes.data, hf.shape[0], hf.shape[1]
        else:
            hf_ptr, rows, cols = None, 0, 0
        seed = int(cfg["basic"].get("seed", 0)) & 0xFFFFFFFFFFFFFFFF
        _lib.check(self._lib.b200_t1_create(C.byref(self._c_model), C.byref(self._c_cfg), hf_ptr, rows, cols, self.num_envs,
                                            self.sim_device_id, seed, C.byref(handle)), "b200_t1_create")
        self._h = handle
        self.terrain._handle = handle

        f_rows, i_rows = self._lib.b200_t1_num_float_rows(), self._lib.b200_t1_num_int_rows()
        self._fstate = torch.zeros(f_rows, self.num_envs, dtype=torch.float32, device=dev)
        self._istate = torch.zeros(i_rows, self.num_envs, dtype=torch.int32, device=dev)
        _lib.check(self._lib.b200_t1_bind_state(self._h, self._fstate.data_ptr(), self._istate.data_ptr()), "bind_state")
        self._ffields = _lib.field_table(0)
        self._ifields = _lib.field_table(1)
        # command-curriculum grid (envs/t1.py:255-262): a device tensor the kernels update in place; the Runner reads / assigns
        # `env.curriculum_prob` for checkpoints (utils/runner.py:91,211) - the setter copies INTO this buffer
        cm = cfg["commands"]
        self._curriculum_prob = torch.zeros(1 + 2 * cm["lin_vel_levels"], 1 + 2 * cm["ang_vel_levels"], dtype=torch.float, device=dev)
        self._curriculum_prob[cm["lin_vel_levels"], cm["ang_vel_levels"]] = 1.0
        if cm.get("curriculum"):
            _lib.check(self._lib.b200_t1_bind_curriculum(self._h, self._curriculum_prob.data_ptr(), self._curriculum_prob.shape[0],
                                                         self._curriculum_prob.shape[1]), "bind_curriculum")

        self._get_env_origins()
        _lib.check(self._lib.b200_t1_init_params(self._h, self._env_index_base, self._total_envs, self._stream()), "init_params")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _fview(self, name):
        row, cnt = self._ffields[name]
        v = self._fstate[row:row + cnt]
        return v[0] if cnt == 1 else v.t()

    def _iview(self, name):
        row, cnt = self._ifields[name]
        v = self._istate[row:row + cnt]
        return v[0] if cnt == 1 else v.t()

    def _get_env_origins(self):
        heights = None
        if self.cfg["terrain"]["type"] != "plane":
            def heights(xyz):
                return self.terrain.terrain_heights(torch.from_numpy(np.ascontiguousarray(xyz)).to(self.device)).cpu().numpy()
        origins = config.env_origins(self.cfg, self._total_envs, heights)
        lo = self._env_index_base
        self.env_origins = self._fview("env_origins")
        self.env_origins.copy_(torch.from_numpy(origins[lo:lo + self.num_envs]).to(self.device))

    def _init_buffers(self):
        cfg, dev, n = self.cfg, self.device, self.num_envs
        self.num_obs = cfg["env"]["num_observations"]
        self.num_privileged_obs = cfg["env"]["num_privileged_obs"]
        self.num_actions = cfg["env"]["num_actions"]
        if (self.num_obs, self.num_privileged_obs, self.num_actions) != (_abi.NOBS, _abi.NPRIV, _abi.NU):
            raise ValueError("the T1 kernels are built for 47 observations, 14 privileged observations and 12 actions")
        self.dt = cfg["control"]["decimation"] * cfg["sim"]["dt"]

        self.obs_buf = torch.zeros(n, self.num_obs, dtype=torch.float, device=dev)
        self.privileged_obs_buf = torch.zeros(n, self.num_privileged_obs, dtype=torch.float, device=dev)
        self.rew_buf = torch.zeros(n, dtype=torch.float, device=dev)
        self._done_u8 = torch.ones(n, dtype=torch.uint8, device=dev)
        self._time_outs_u8 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.reset_buf = self._done_u8.view(torch.bool)
        self.extras = {"rew_terms": {}, "privileged_obs": self.privileged_obs_buf, "time_outs": self._time_outs_u8.view(torch.bool)}
        self.common_step_counter = 0

        # reference-named views into the structure-of-arrays state
        for name in ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel",
                     "last_dof_targets", "torques", "commands", "gait_frequency", "gait_process", "base_lin_vel",
                     "base_ang_vel", "projected_gravity", "filtered_lin_vel", "filtered_ang_vel", "feet_roll", "feet_yaw",
                     "dof_stiffness", "dof_damping", "dof_friction", "base_mass_scaled"):
            setattr(self, name, self._fview(name))
        self.base_pos = self.root_states[:, 0:3]
        self.base_quat = self.root_states[:, 3:7]
        self.feet_pos = self._fview("feet_pos").unflatten(1, (2, 3))
        self.feet_quat = self._fview("feet_quat").unflatten(1, (2, 4))
        self.last_feet_pos = self._fview("last_feet_pos").unflatten(1, (2, 3))
        self.pushing_forces = self._fview("pushing_forces")   # [N,3] on the Trunk (the reference keeps [N,13,3], 12 rows always 0)
        self.pushing_torques = self._fview("pushing_torques")
        self.episode_length_buf = self._iview("episode_length_buf")
        self.cmd_resample_time = self._iview("cmd_resample_time")
        self.delay_steps = self._iview("delay_steps")
        self.feet_contact = self._iview("feet_contact")
        self.time_out_buf = self._iview("time_out_buf")
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0], device=dev).repeat((n, 1))
        self.default_dof_pos = torch.tensor([list(self._c_cfg.default_dof_pos)], dtype=torch.float, device=dev)

        self.env_curriculum_level = self._iview("env_curriculum_level")   # [N, 2] (lin, ang) level of the current command

    # ---- command curriculum surface (envs/t1.py:255-262, 419-422; read by utils/runner.py:91,198-201,211) -------------------
    @property
    def curriculum_prob(self):
        return self._curriculum_prob

    @curriculum_prob.setter
    def curriculum_prob(self, value):
        self._curriculum_prob.copy_(torch.as_tensor(value).to(device=self._curriculum_prob.device, dtype=torch.float).reshape(self._curriculum_prob.shape))

    @property
    def mean_lin_vel_level(self):
        return torch.mean(torch.abs(self.env_curriculum_level[:, 0]).float())

    @property
    def mean_ang_vel_level(self):
        return torch.mean(torch.abs(self.env_curriculum_level[:, 1]).float())

    @property
    def max_lin_vel_level(self):
        return torch.max(torch.abs(self.env_curriculum_level[:, 0]))

    @property
    def max_ang_vel_level(self):
        return torch.max(torch.abs(self.env_curriculum_level[:, 1]))

    def _prepare_reward_function(self):
        terms = config.reward_terms(self.cfg)
        self.reward_names = [name for name, _ in terms]
        self.reward_scales = {name: scale for name, scale in terms}
        self._rew_terms = torch.zeros(max(1, len(terms)), self.num_envs, dtype=torch.float, device=self.device)
        self.extras["rew_terms"] = {name: self._rew_terms[k] for k, name in enumerate(self.reward_names)}

    # ---- the task API ------------------------------------------------------------------------------------------------
    def reset(self):
        """Reset all robots"""
        _lib.check(self._lib.b200_t1_reset(self._h, self.obs_buf.data_ptr(), self.privileged_obs_buf.data_ptr(), self._stream()),
                   "b200_t1_reset")
        return self.obs_buf, self.extras

    def step(self, actions, _device_counter=False, _out=None):
        """`_out` (not in the reference surface): (obs, privileged_obs, rew, done_u8) tensors the step writes INSTEAD of the env's own
        obs_buf / privileged_obs_buf / rew_buf / reset_buf - the Runner passes the rows of its rollout storage, which removes four
        copy kernels per step.  The returned tensors are then those; the env's own buffers keep their previous content."""
        a = actions
        if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.obs_buf.device:
            a = a.to(device=self.obs_buf.device, dtype=torch.float32).contiguous()
        self.common_step_counter += 1
        obs, priv, rew, done = (self.obs_buf, self.privileged_obs_buf, self.rew_buf, self._done_u8) if _out is None else _out
        _lib.check(self._lib.b200_t1_step(self._h, a.data_ptr(), obs.data_ptr(), priv.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                          self._time_outs_u8.data_ptr(), self._rew_terms.data_ptr(),
                                          -1 if _device_counter else self.common_step_counter, self._stream()), "b200_t1_step")
        self.render()
        if _out is None:
            return self.obs_buf, self.rew_buf, self.reset_buf, self.extras
        return obs, rew, done.view(torch.bool), {"rew_terms": self.extras["rew_terms"], "privileged_obs": priv, "time_outs": self.extras["time_outs"]}

    # ---- extras (not in the reference surface) ---------------------------------------------------------------------
    def physics(self, actions_or_torques, n_substeps, apply_pd=True, qacc_out=None):
        """the decimated PD-torque loop alone (envs/t1.py:439-456); BASELINE config 5 and the physics parity tests"""
        a = actions_or_torques.to(device=self.obs_buf.device, dtype=torch.float32).contiguous()
        _lib.check(self._lib.b200_t1_physics(self._h, a.data_ptr(), int(n_substeps), 1 if apply_pd else 0,
                                             qacc_out.data_ptr() if qacc_out is not None else None, self._stream()), "b200_t1_physics")

    def post_physics(self, noise=True):
        self.common_step_counter += 1
        _lib.check(self._lib.b200_t1_post_physics(self._h, self.obs_buf.data_ptr(), self.privileged_obs_buf.data_ptr(),
                                                  self.rew_buf.data_ptr(), self._done_u8.data_ptr(), self._time_outs_u8.data_ptr(),
                                                  self._rew_terms.data_ptr(), self.common_step_counter, 1 if noise else 0,
                                                  self._stream()), "b200_t1_post_physics")
        return self.obs_buf, self.rew_buf, self.reset_buf, self.extras

    def episode_stats(self):
        """device-side episode statistics (replaces utils/recorder.py:36-62): dict of means over the episodes that
        finished since the last call, and their count. Synchronises the stream."""
        n = len(self.reward_names)
        sums = (C.c_double * (n + 2))()
        cnt = C.c_int64()
        _lib.check(self._lib.b200_t1_episode_stats(self._h, sums, C.byref(cnt), self._stream()), "episode_stats")
        k = max(1, cnt.value)
        out = {"reward": sums[0] / k, "steps": sums[n + 1] / k}
        for i, name in enumerate(self.reward_names):
            out[name] = sums[1 + i] / k
        return out, cnt.value

    def episode_stats_async(self, out_pinned):
        """stream-ordered read-and-clear of the same accumulators into a pinned float64 tensor of n_rew + 3 elements, no
        synchronisation (decode with `decode_episode_stats` once an event recorded after this call has completed)"""
        assert out_pinned.is_pinned() and out_pinned.dtype == torch.float64 and out_pinned.numel() >= len(self.reward_names) + 3
        _lib.check(self._lib.b200_t1_episode_stats_async(self._h, out_pinned.data_ptr(), self._stream()), "episode_stats_async")

    def decode_episode_stats(self, host):
        n = len(self.reward_names)
        cnt = int(host[n + 2].item() + 0.5)
        k = max(1, cnt)
        out = {"reward": host[0].item() / k, "steps": host[n + 1].item() / k}
        for i, name in enumerate(self.reward_names):
            out[name] = host[1 + i].item() / k
        return out, cnt

    def inject_rng(self, table):
        """parity-test hook: int32/uint32 tensor [slots, 12, N] on the device (or None to restore the Philox draws)"""
        self._inject = table
        _lib.check(self._lib.b200_t1_inject_rng(self._h, table.data_ptr() if table is not None else None), "inject_rng")

    def rng_samples(self, step, purpose, sub, kind):
        out = torch.empty(4, self.num_envs, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.b200_rng_fill(self._h, int(step), int(purpose), int(sub), int(kind), out.data_ptr(), self._stream()), "rng_fill")
        return out

    def counters(self):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self._lib.b200_t1_counters(self._h, C.byref(a), C.byref(b), self._stream()), "counters")
        return a.value, b.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200_t1_destroy(self._h)
            self._h = None
            self.terrain._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
