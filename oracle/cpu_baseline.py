"""CPU baseline of the T1 training iteration, timed on the host cores.  TEST / BENCH INFRASTRUCTURE ONLY.

Neither reference physics engine can run here or on the GPU box (Isaac Gym: closed binary; mujoco: not installed, no
network), and the reference's Python is not on the GPU box, so the baseline is this repo's CPU PORT of the path
("kind": "port"): the FP64 C restatement of the MuJoCo smooth pipeline (oracle/physics_oracle.c, one env per OpenMP
thread) for the decimated physics loop, and the torch-CPU restatement of utils/model.py + utils/runner.py:123-180
(oracle/learner.py, pinned to the live reference by tests/golden) for policy inference and the PPO update.
Each call times a BOUNDED SAMPLE of one iteration and extrapolates linearly in the number of envs / epochs.
"""
import ctypes as C
import os
import time

import numpy as np
import torch


def _physics_sample(n_envs, n_steps, decimation=10, seed=0):
    """seconds for n_steps env steps (x decimation ticks) of n_envs envs, all host cores (OpenMP)"""
    from booster_gym_b200 import robot
    from oracle import physics as op

    md = robot.model_d()
    lib = op.port_lib()
    rng = np.random.default_rng(seed)
    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2, dtype=np.float64)
    envs = (op.Env * n_envs)(*[op.make_env(md, pos=(0.0, 0.0, 0.68), q=q0 + rng.normal(0, 0.05, 12)) for _ in range(n_envs)])
    kp = np.tile(np.array([200.0, 200, 200, 200, 50, 50] * 2), (n_envs, 1))
    kd = np.tile(np.array([5.0, 5, 5, 5, 1, 1] * 2), (n_envs, 1))
    fr = np.zeros((n_envs, 12))
    lim = np.array([45.0, 30, 30, 60, 24, 15] * 2)
    delay = np.zeros(n_envs, dtype=np.int32)
    lt = np.tile(q0, (n_envs, 1))
    pf = np.zeros((n_envs, 3))
    pt = np.zeros((n_envs, 3))
    tm = np.zeros((n_envs, 12))
    terr = op.make_terrain()
    P = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    t0 = time.perf_counter()
    for _ in range(n_steps):
        act = np.clip(rng.normal(0, 0.3, (n_envs, 12)), -1, 1)
        lib.t1p_env_physics(C.byref(md), envs, n_envs, P(act), P(q0), C.c_double(1.0), P(kp), P(kd), P(fr), P(lim), P(delay),
                            P(lt), decimation, P(tm))
    return time.perf_counter() - t0


def iteration_sample(num_envs=4096, horizon=24, epochs=20, phys_envs=None, phys_steps=4, ppo_epochs=1, threads=None):
    """Time a bounded sample of one training iteration on the CPU and extrapolate.
    Returns dict(seconds_per_iteration, env_steps_per_s, cores, sample, parts)."""
    from oracle import learner as L

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    phys_envs = phys_envs or max(cores * 32, 256)
    t_phys = _physics_sample(phys_envs, phys_steps)
    phys_per_env_step = t_phys / (phys_envs * phys_steps)
    sd = L.init_params(0)
    buf, last_obs, last_priv = L.synthetic_rollout(horizon, num_envs, seed=0)
    t0 = time.perf_counter()
    with torch.no_grad():
        for t in range(2):
            mu = L.actor_mean(sd, buf["obses"][t])
            _ = mu + torch.exp(sd["logstd"]) * torch.randn_like(mu)
    t_act = (time.perf_counter() - t0) / 2
    omu, osig, olp = L.old_dist(sd, buf["obses"], buf["actions"])
    adam = L.new_adam(sd)
    L.epoch(sd, adam, buf, last_obs, last_priv, omu, osig, olp, 1e-5)  # untimed warm-up: allocator, thread pool, page faults
    t0 = time.perf_counter()
    for _ in range(ppo_epochs):
        L.epoch(sd, adam, buf, last_obs, last_priv, omu, osig, olp, 1e-5)
    t_epoch = (time.perf_counter() - t0) / ppo_epochs
    rollout = horizon * (num_envs * phys_per_env_step + t_act)
    update = epochs * t_epoch
    total = rollout + update
    return dict(
        seconds_per_iteration=total,
        env_steps_per_s=num_envs * horizon / total,
        cores=cores,
        sample=(f"physics: {phys_envs} envs x {phys_steps} env-steps (x10 ticks) of the FP64 -O3 host build of the CRBA/RNE/LTDL recursion (oracle/physics_port.cpp), OpenMP {cores} threads, "
                f"scaled to {num_envs} envs x {horizon} steps; policy: 2 x actor({num_envs}) torch-CPU; update: {ppo_epochs} (after 1 untimed warm-up) of "
                f"{epochs} full-batch epochs (T={horizon}, N={num_envs}) torch-CPU autograd, scaled x{epochs // ppo_epochs}; "
                f"obs/reward pass not included"),
        parts=dict(physics_s_per_env_step=phys_per_env_step, policy_s_per_step=t_act, epoch_s=t_epoch,
                   rollout_s=rollout, update_s=update),
    )


if __name__ == "__main__":
    import json

    print(json.dumps(iteration_sample(), indent=1))
