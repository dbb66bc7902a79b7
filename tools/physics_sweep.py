#!/usr/bin/env python3
"""BASELINE.json configs[4]: physics-only sweep of k_physics over N envs per GPU, 1000 substeps,
(i) contact-free (base held around z = 2 m, PD to the default pose) and (ii) full foot contact (settled standing pose), next to
the CPU port (oracle/physics_port.cpp, all host cores).  Writes one JSON document.

    python tools/physics_sweep.py [--out gpurun_out/physics_sweep.json] [--sizes 1024,4096,...]

FLOP figure: 1.28e4 per env-substep (FMA = 2).  FP32 ceiling: 148 SMs x 128 lanes x 2 x SM clock.
"""
import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

FLOP_PER_ENV_SUBSTEP = 1.28e4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/physics_sweep.json")
    ap.add_argument("--sizes", default="1024,4096,16384,65536,262144")
    ap.add_argument("--substeps", type=int, default=1000)
    args = ap.parse_args()
    import numpy as np
    import torch
    import yaml

    from booster_gym_b200.envs import T1

    base = yaml.load(open(os.path.join(ROOT, "envs", "T1.yaml")).read(), Loader=yaml.FullLoader)
    dev = torch.device("cuda:0")
    sm_mhz = 1965.0
    try:
        import subprocess

        sm_mhz = float(subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"]).decode().split()[0])
    except Exception:
        pass
    peak = 148 * 128 * 2 * sm_mhz * 1e6
    rows = []
    for n in [int(x) for x in args.sizes.split(",")]:
        cfg = copy.deepcopy(base)
        cfg["env"]["num_envs"] = n
        cfg["terrain"]["type"] = "plane"
        np.random.seed(42)
        env = T1(cfg)
        zero = torch.zeros(n, 12, device=dev)
        for scenario in ("contact_free", "foot_contact"):
            env.reset()
            # nominal pose for every env (the reset distribution adds joint / yaw noise under which an uncontrolled robot falls within
            # ~1 s; only the soles carry contact geometry, so a fallen robot is not a "full foot contact" state)
            env.dof_pos[:] = env.default_dof_pos
            env.dof_vel[:] = 0.0
            env.root_states[:, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0], device=dev)
            env.root_states[:, 7:13] = 0.0
            if scenario == "contact_free":
                env.root_states[:, 2] = 2.0
            else:
                env.root_states[:, 2] = 0.68
                for _ in range(25):          # settle on the feet: 250 ticks = 0.5 s
                    env.physics(zero, 10)
            torch.cuda.synchronize()
            snap = env._fstate.clone()       # restored every 100 substeps so that the whole timed window stays in the scenario's regime
            z_keep = env.root_states[:, 2].clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for rep in range(2):             # rep 0 = warm-up
                e0.record()
                for i in range(args.substeps // 10):
                    env.physics(zero, 10)
                    if i % 10 == 9:
                        if scenario == "contact_free":   # hold the base near z = 2 m: never reaches the ground
                            env.root_states[:, 2] = z_keep
                            env.root_states[:, 7:10] = 0.0
                        else:
                            if i == args.substeps // 10 - 1:
                                standing = float((env.root_states[:, 2] > 0.6).float().mean().item())
                            env._fstate.copy_(snap)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            rate = n * args.substeps / (ms * 1e-3)
            assert torch.isfinite(env.root_states).all()
            if scenario == "contact_free":
                standing = None
            rows.append({"num_envs": n, "scenario": scenario, "substeps": args.substeps, "ms": ms, "env_substeps_per_s": rate,
                         "env_steps_per_s_decimation10": rate / 10, "gflops_algorithmic": rate * FLOP_PER_ENV_SUBSTEP / 1e9,
                         "frac_fp32_peak": rate * FLOP_PER_ENV_SUBSTEP / peak, "mean_base_z": float(env.root_states[:, 2].mean().item()),
                         "frac_standing_before_last_restore": standing})
            print(rows[-1], flush=True)
        env.close()
        del env
        torch.cuda.empty_cache()
    cpu = None
    try:
        from oracle import cpu_baseline

        cpu_baseline._physics_sample(64, 1)
        t = cpu_baseline._physics_sample(512, 4)
        cpu = {"env_substeps_per_s": 512 * 4 * 10 / t, "cores": os.cpu_count(), "kind": "port",
               "sample": "512 envs x 4 env-steps x 10 ticks, FP64 -O3 host build of the same recursion (oracle/physics_port.cpp), OpenMP all cores"}
    except Exception as ex:
        cpu = {"error": str(ex)}
    doc = {"workload": "BASELINE configs[4]: physics-only k_physics sweep, %d substeps per size" % args.substeps, "sm_max_mhz": sm_mhz,
           "fp32_peak_tflops": peak / 1e12, "flop_per_env_substep": FLOP_PER_ENV_SUBSTEP, "rows": rows, "cpu_baseline": cpu}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
