"""BASELINE configs[0]: play_mujoco-style closed loop (play_mujoco.py:717-764) - 1 policy step = obs -> random-init actor
(deterministic dist.loc) -> PD targets -> 10 physics ticks - run through the PUBLIC API on the GPU (T1.step + Learner.act)
and, independently, on the CPU with the FP64 physics oracle + the torch actor oracle.  The reference trajectory is the CPU
oracle (MuJoCo itself is not installable: parity unpinned); contact makes the system chaotic, so the
claim is bounded short-horizon drift (stated in the asserts at the end; measured: micrometres over the first 0.6 s)."""
import copy
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_closed_loop_rollout_tracks_fp64_oracle(t1_cfg):
    from booster_gym_b200 import robot
    from booster_gym_b200.envs import T1
    from booster_gym_b200.learner import Learner
    from oracle import learner as L
    from oracle import physics as op
    from test_gpu_env import oracle_envs

    cfg = copy.deepcopy(t1_cfg)
    n = 8
    cfg["env"]["num_envs"] = n
    cfg["terrain"]["type"] = "plane"
    cfg.setdefault("asset", {})["effort_limits"] = "mjcf"   # play_mujoco.py clips to the MJCF ctrlrange (SURVEY 8a note 10), not the URDF efforts
    cfg["noise"] = {}
    for k in list(cfg["randomization"].keys()):
        if isinstance(cfg["randomization"][k], dict):
            cfg["randomization"][k] = None
    np.random.seed(0)
    env = T1(cfg)
    lrn = Learner(cfg, n, "cuda:0")
    torch.manual_seed(42)
    sd = L.init_params(42)
    lrn.load_state_dict(sd)
    sd64 = {k: v.double() for k, v in sd.items()}
    env.reset()
    env.commands[:] = torch.tensor([0.5, 0.0, 0.0], device="cuda")
    env.gait_frequency[:] = 1.5
    env.cmd_resample_time[:] = 1000000
    env.delay_steps[:] = 0
    torch.cuda.synchronize()
    md, oenvs = oracle_envs(env, range(n))
    arr = (op.Env * n)(*oenvs)
    q0 = env.default_dof_pos.cpu().double().numpy().ravel()
    kp = np.tile(np.array(list(env._c_cfg.kp_nominal), dtype=np.float64), (n, 1))
    kd = np.tile(np.array(list(env._c_cfg.kd_nominal), dtype=np.float64), (n, 1))
    fr = np.zeros((n, 12)); lim = env.torque_limits.cpu().double().numpy(); delay = np.zeros(n, np.int32)
    assert lim.ravel()[1] == 45.0 and lim.ravel()[3] == 65.0      # hip roll / knee: ctrlrange, not the URDF's 30 / 60
    lt = env.last_dof_targets.cpu().double().numpy().copy()
    pf = np.zeros((n, 3)); pt = np.zeros((n, 3)); tm = np.zeros((n, 12)); terr = op.make_terrain()
    P = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib = op.lib()

    def oracle_obs(gp, last_act):
        rows = []
        for e in range(n):
            x, y, z, w = arr[e].quat[:]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                          [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
            pg = R.T @ np.array([0, 0, -1.0])
            rows.append(np.concatenate([pg, np.array(arr[e].wb[:]), [0.5, 0.0, 0.0], [np.cos(2 * np.pi * gp), np.sin(2 * np.pi * gp)],
                                        np.array(arr[e].q[:]) - q0, 0.1 * np.array(arr[e].qd[:]), last_act[e]]))
        return np.stack(rows)

    # first observation from the post-reset state (play_mujoco.py:734-744), identical on both sides
    gp = float(env.gait_process[0].item())
    last_act = np.zeros((n, 12))
    obs_o = oracle_obs(gp, last_act)
    obs_g = torch.from_numpy(obs_o.astype(np.float32)).cuda()
    act_g = torch.empty(n, 12, device="cuda")
    drift = []
    for step in range(50):   # 500 physics ticks (BASELINE configs[0])
        lrn.act(obs_g, act_g, deterministic=True)
        obs_g, rew, done, _ = env.step(act_g)
        mu = L.actor_mean(sd64, torch.from_numpy(obs_o)).numpy()
        a = np.clip(mu, -1, 1).copy()
        bad = lib.t1o_env_physics(C.byref(md), arr, n, P(a), P(q0), C.c_double(1.0), P(kp), P(kd), P(fr), P(lim), P(delay), P(lt),
                                  P(pf), P(pt), C.byref(terr), 10, P(tm), 0)
        assert bad == 0
        gp = float(np.fmod(gp + 0.02 * 1.5, 1.0))
        last_act = a
        obs_o = oracle_obs(gp, last_act)
        torch.cuda.synchronize()
        if bool(done.any()):
            break
        rs = env.root_states.cpu().double().numpy(); q = env.dof_pos.cpu().double().numpy()
        dpos = max(np.abs(rs[e, 0:3] - np.array(arr[e].pos[:])).max() for e in range(n))
        dq = max(np.abs(q[e] - np.array(arr[e].q[:])).max() for e in range(n))
        dobs = np.abs(obs_g.cpu().double().numpy() - obs_o).max()
        drift.append((step, dpos, dq, dobs, float(rs[:, 2].min())))
    for row in drift:
        print("step %2d  |dpos| %.2e  |dq| %.2e  |dobs| %.2e  z_min %.3f" % row)
    assert len(drift) >= 30, "the random-init policy should keep the robot above the termination height for 0.6 s"
    # measured on B200: |dpos| <= 2e-6 m, |dq| <= 2e-5 rad over the first 0.6 s (touch-down at step 5); stated bound:
    for row in drift[:30]:
        assert row[1] < 1e-4 and row[2] < 1e-3 and row[3] < 1e-3, row
    # afterwards the statically unstable stance (ankle kp 50) starts to fall and the chaotic growth sets in: bounded drift
    for row in drift[30:]:
        assert row[1] < 5e-2 and row[2] < 2e-1, row
