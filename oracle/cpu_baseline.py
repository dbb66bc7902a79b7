"""CPU baseline of the T1 training iteration, timed on the host cores.  TEST / BENCH INFRASTRUCTURE ONLY.

Neither reference physics engine can run here or on the GPU box (Isaac Gym: closed binary; mujoco: not installed, no
network), and the reference's Python is not on the GPU box, so the baseline is this repo's CPU PORT of the path
("kind": "port"): the FP64 -O3 host build of the rigid-body recursion (oracle/physics_port.cpp, one env per OpenMP thread)
for the decimated physics loop, the numpy restatement of the post-physics step - observations, 23 rewards, termination,
resets (oracle/env_oracle.py, pinned to the reference by tests/golden/env_step_*.npz), and the torch-CPU restatement of
utils/model.py + utils/runner.py:123-180 (oracle/learner.py, pinned by tests/golden/learner_*.npz) for policy inference and
the PPO update.

One call of `iteration()` runs ONE WHOLE training iteration - 24 x [policy + sampling, 10-tick physics, post-physics pass],
the old-distribution pass and all 20 full-batch epochs - at a REDUCED number of environments (the bounded sample: default
1024 of the workload's 4096 per GPU), and reports the env-steps/s of exactly that run; nothing is extrapolated.  CPU cost is
linear in the number of environments for every part, so the rate carries over to the full size.
"""
import copy
import ctypes as C
import os
import time

import numpy as np
import torch


class CpuIteration:
    def __init__(self, num_envs=1024, horizon=24, epochs=20, threads=None, seed=0):
        import yaml

        from booster_gym_b200 import robot
        from oracle import learner as L
        from oracle import physics as op
        from oracle.env_oracle import EnvOracle
        from oracle.ref_harness import make_table
        from oracle.synth import synthetic_state

        self.L, self.op = L, op
        self.n, self.T, self.E = num_envs, horizon, epochs
        self.cores = threads or os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        cfg = yaml.safe_load(open(os.path.join(root, "envs", "T1.yaml")))
        cfg = copy.deepcopy(cfg)
        cfg["terrain"]["type"] = "plane"
        import json

        mj = json.load(open(os.path.join(root, "booster_gym_b200", "assets", "t1_model.json")))
        self.env = EnvOracle(cfg, synthetic_state(num_envs, seed, False, ep_case=False), None, mj)
        self.table = make_table(num_envs, seed + 1)
        self.md = robot.model_d()
        self.lib = op.port_lib()
        rng = np.random.default_rng(seed)
        self.rng = rng
        n = num_envs
        self.q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2, dtype=np.float64)
        self.envs = (op.Env * n)(*[op.make_env(self.md, pos=(0.0, 0.0, 0.68), q=self.q0 + rng.normal(0, 0.05, 12)) for _ in range(n)])
        self.kp = np.tile(np.array([200.0, 200, 200, 200, 50, 50] * 2), (n, 1))
        self.kd = np.tile(np.array([5.0, 5, 5, 5, 1, 1] * 2), (n, 1))
        self.fr = np.zeros((n, 12))
        self.lim = np.array([45.0, 30, 30, 60, 24, 15] * 2)
        self.delay = np.zeros(n, dtype=np.int32)
        self.lt = np.tile(self.q0, (n, 1))
        self.tm = np.zeros((n, 12))
        self.sd = L.init_params(0)
        self.buf, self.last_obs, self.last_priv = L.synthetic_rollout(horizon, num_envs, seed=0)
        self.adam = L.new_adam(self.sd)
        self.step = 1

    def iteration(self):
        """one whole iteration; returns (seconds, parts)"""
        L, lib, md = self.L, self.lib, self.md
        P = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        parts = {"policy": 0.0, "physics": 0.0, "post_physics": 0.0, "update": 0.0}
        t_start = time.perf_counter()
        for t in range(self.T):
            t0 = time.perf_counter()
            with torch.no_grad():
                mu = L.actor_mean(self.sd, self.buf["obses"][t])
                act = (mu + torch.exp(self.sd["logstd"]) * torch.randn_like(mu)).clamp(-1, 1).double().numpy()
            t1 = time.perf_counter()
            lib.t1p_env_physics(C.byref(md), self.envs, self.n, P(np.ascontiguousarray(act)), P(self.q0), C.c_double(1.0), P(self.kp),
                                P(self.kd), P(self.fr), P(self.lim), P(self.delay), P(self.lt), 10, P(self.tm))
            t2 = time.perf_counter()
            self.env.step_post(self.table, self.step)   # observations, rewards, termination, resets (numpy, one core)
            self.step += 1
            t3 = time.perf_counter()
            parts["policy"] += t1 - t0
            parts["physics"] += t2 - t1
            parts["post_physics"] += t3 - t2
        t0 = time.perf_counter()
        omu, osig, olp = L.old_dist(self.sd, self.buf["obses"], self.buf["actions"])
        lr = 1e-5
        for _ in range(self.E):
            o = L.epoch(self.sd, self.adam, self.buf, self.last_obs, self.last_priv, omu, osig, olp, lr)
            lr = o["lr"]
        parts["update"] = time.perf_counter() - t0
        return time.perf_counter() - t_start, parts

    def sample_text(self):
        return (f"one WHOLE iteration at {self.n} envs (24 x [torch-CPU actor + sampling, FP64 -O3 OpenMP port of the 10-tick physics loop, numpy "
                f"post-physics pass (obs / 23 rewards / termination / resets)] + old-dist pass + {self.E} full-batch torch-CPU autograd epochs on "
                f"{self.T * self.n} samples), {self.cores} threads; rate = {self.n} x {self.T} env-steps / measured seconds, nothing extrapolated")


def iteration_sample(num_envs=1024, iterations=1, warmup=1, threads=None):
    """`warmup` untimed + `iterations` timed whole iterations at num_envs environments."""
    it = CpuIteration(num_envs=num_envs, threads=threads)
    for _ in range(warmup):
        it.iteration()
    secs, parts = [], None
    for _ in range(iterations):
        s, parts = it.iteration()
        secs.append(s)
    sec = sum(secs) / len(secs)
    return dict(seconds_per_iteration=sec, seconds=secs, env_steps_per_s=num_envs * it.T / sec, cores=it.cores, sample=it.sample_text(),
                parts=parts, num_envs=num_envs)


if __name__ == "__main__":
    import json

    print(json.dumps(iteration_sample(), indent=1))
