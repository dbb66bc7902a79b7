#!/usr/bin/env python3
"""bench.py - BASELINE.json metric on BASELINE.json configs[1]: T1 flat terrain, 4096 envs per GPU, one "step" =
one PPO iteration = 24-step rollout (policy inference + sampling + env.step + rollout storage) + the PPO update
(old-dist pass + 20 full-batch epochs of values/GAE/losses/backward/clip/Adam/KL-lr).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--num-envs 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = env-steps/s of the whole job over the FULL iteration (rollout + update), states
resident in HBM; `e2e` = the same loop driven through the public Runner API with the step's synthetic state batch
uploaded from pinned host memory and the iteration's scalars read back, inside the timed region.  Extra keys give the
rollout-only rate (the north_star's "env-steps/s incl. policy") and the PPO update time.
`--impl reference` times the CPU port of the same iteration (oracle/cpu_baseline.py) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.chdir(ROOT)
sys.path.insert(0, ROOT)

METRIC = "T1 env-steps/sec incl. policy (full PPO iteration: 24-step rollout + 20-epoch update)"
UNIT = "env-steps/s"
WORKLOAD = "configs[1]: T1 flat terrain, 4096 envs per GPU, 24-step rollout + PPO update (20 full-batch epochs), seed 42"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200")
    ap.add_argument("--num-envs", type=int, default=None, help="envs per GPU (default: 4096 for --config 1, 8192 for --config 2)")
    ap.add_argument("--config", type=int, default=1, choices=(1, 2),
                    help="1 = BASELINE.json configs[1] (flat terrain, 4096 envs/GPU; the bench line); 2 = configs[2]/[3] (rough heightfield "
                         "terrain + kicks/pushes + domain randomisation, 8192 envs/GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graphs", type=int, default=1, help="replay the rollout from a CUDA graph (Runner.train's default); 0 = eager launches")
    args = ap.parse_args()
    if args.num_envs is None:
        args.num_envs = 4096 if args.config == 1 else 8192
    return args


def workload(args):
    if args.config == 1:
        return WORKLOAD.replace("4096 envs", f"{args.num_envs} envs")
    return (f"configs[2]: T1 rough heightfield terrain (trimesh: 4 random-uniform + 4 discrete-obstacle tiles) + kicks/pushes + domain "
            f"randomisation, {args.num_envs} envs per GPU, 24-step rollout + PPO update (20 full-batch epochs), seed 42")


def config_dict(args, world):
    return {"workload": workload(args), "num_envs_per_gpu": args.num_envs, "total_envs": args.num_envs * world, "horizon": 24,
            "mini_epochs": 20, "terrain": "plane" if args.config == 1 else "trimesh", "parallelism": f"env-sharded dp{world}",
            "l2": "working set (the learner streams ~1.9 GB of fp32 activations per epoch through a >400 MB workspace per GPU) is larger than "
                  "the 126 MB L2; no flush needed",
            "rollout_cuda_graph": bool(args.graphs)}


# ---------------------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """CPU port of the iteration on the host cores; rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline

    vals, parts = [], None
    for _ in range(max(1, args.warmup if args.warmup < 1 else 0)):
        pass
    t_all = time.perf_counter()
    for k in range(max(1, args.steps)):
        r = cpu_baseline.iteration_sample(num_envs=args.num_envs, phys_envs=256, phys_steps=2, ppo_epochs=1)
        vals.append(r)
        if time.perf_counter() - t_all > 150:
            break
    sec = sum(v["seconds_per_iteration"] for v in vals) / len(vals)
    value = args.num_envs * 24 / sec
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": 0,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference", "config": config_dict(args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": vals[0]["cores"], "kind": "port", "sample": vals[0]["sample"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "one CPU process on rank 0 regardless of --gpus; each step is a bounded sample extrapolated to a full iteration",
           "parts": vals[0]["parts"]}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def b200_arm(args):
    import ctypes as C

    import torch

    from booster_gym_b200 import _abi, _lib
    from booster_gym_b200.utils.runner import Runner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    lib = _lib.load()
    runner = Runner(test=False, argv=["--task", "T1", "--num_envs", str(args.num_envs), "--headless", "True"],
                    cfg_overrides={"terrain": {"type": "plane"}, "runner": {"use_wandb": False}} if args.config == 1 else
                    {"terrain": {"type": "trimesh"}, "runner": {"use_wandb": False}})
    dev = torch.device(runner.device)
    env, lrn = runner.env, runner.learner
    T, N, E = runner.cfg["runner"]["horizon_length"], env.num_envs, runner.cfg["runner"]["mini_epochs"]
    stream = torch.cuda.current_stream(dev)

    obs, infos = env.reset()
    priv = infos["privileged_obs"]

    graph, graph_launches = None, 0
    if args.graphs:
        # warm the rollout once on a side stream, then capture the 24 steps as one graph
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            runner.rollout(obs, priv)
        torch.cuda.current_stream(dev).wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        lc0 = lib.b200_launch_count()
        with torch.cuda.graph(graph):
            runner.rollout(obs, priv)
        graph_launches = lib.b200_launch_count() - lc0   # kernels of this library inside one replay (the host counter only sees the capture)

    def iteration():
        if graph is not None:
            graph.replay()
        else:
            runner.rollout(obs, priv)
        runner.update(obs, priv)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(3, args.warmup)):
        iteration()
    # ---- main measurement: states resident in HBM --------------------------------------------------------------
    clocks = ClockSampler(local)
    l0 = lib.b200_launch_count()
    ms_total = timed(iteration, args.steps)
    launches = lib.b200_launch_count() - l0 + (graph_launches * args.steps if graph is not None else 0)
    clock_info = clocks.stop()
    ms_step = ms_total / args.steps
    value = world * N * T / (ms_step * 1e-3)

    # ---- phase split (rollout only / update only), informational ----------------------------------------------------
    ms_roll = timed(lambda: (graph.replay() if graph is not None else runner.rollout(obs, priv)), max(2, args.steps)) / max(2, args.steps)
    ms_upd = timed(lambda: runner.update(obs, priv), max(2, args.steps)) / max(2, args.steps)

    # ---- end to end through the public API with host buffers ----------------------------------------------------
    f_host = env._fstate.cpu().pin_memory()
    i_host = env._istate.cpu().pin_memory()
    sc_host = torch.empty(_abi.SC["COUNT"], dtype=torch.float32).pin_memory()
    ob_host = torch.empty(N, 47, dtype=torch.float32).pin_memory()
    h2d = f_host.numel() * 4 + i_host.numel() * 4
    d2h = sc_host.numel() * 4 + ob_host.numel() * 4

    def e2e_iteration():
        env._fstate.copy_(f_host, non_blocking=True)   # the step's synthetic state batch: pinned host -> HBM
        env._istate.copy_(i_host, non_blocking=True)
        iteration()
        sc_host.copy_(lrn.scalars, non_blocking=True)  # losses / KL / lr of the iteration -> host
        ob_host.copy_(obs, non_blocking=True)
        torch.cuda.synchronize(dev)                    # the host needs the scalars before it may log / continue

    e2e_iteration()
    ms_e2e = timed(e2e_iteration, args.steps) / args.steps
    e2e_value = world * N * T / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel family: CUDA events around every GEMM launch (recorded inside the library on
    # the launching stream), one extra iteration; the family with the largest summed duration is reported
    lib.b200_profile_gemm(1)
    iteration()
    fams = {}
    names = {0: "k_gemm3x (mma.sync 3xTF32: b200_critic_value only)",
             1: "k_tc_rowmajor (tcgen05/TMEM/TMA 3xTF32 with in-smem operand split: MLP forward + dgrad, fused bias/ELU/ELU' epilogue)",
             2: "k_tc_wgrad (tcgen05/TMEM/TMA 3xTF32, MN-major operands, in-smem split: MLP weight gradients)"}
    fam_bytes = {}
    for kind in (0, 1, 2):
        ms_g, fl_g, n_g, by_g = C.c_double(), C.c_double(), C.c_int(), C.c_double()
        lib.b200_profile_gemm_read(kind, C.byref(ms_g), C.byref(fl_g), C.byref(n_g))
        lib.b200_profile_gemm_bytes(kind, C.byref(by_g))
        fams[kind] = (ms_g.value, fl_g.value, n_g.value)
        fam_bytes[kind] = by_g.value
    lib.b200_profile_gemm(0)
    top = max(fams, key=lambda k: fams[k][0])
    ms_top, fl_top, n_top = fams[top]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = fl_top / (ms_top * 1e-3) / 1e12 if ms_top > 0 else 0.0
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch of that family, from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(str(top))
    except Exception:
        pass
    # the family is bounded by whichever roofline it sits closer to: HBM (fp32 activations in / out, 4 B per element) or the
    # tensor pipe (measured dense BF16 peak; the 3xTF32 split executes 3x the algorithmic FLOPs at half the BF16 rate)
    peak_hbm = peaks.get("hbm_gbs", 6500.0)
    gbs = fam_bytes[top] / (ms_top * 1e-3) / 1e9 if ms_top > 0 else 0.0
    hbm_bound = gbs / peak_hbm >= achieved / peak_tf
    roofline = {"kernel": names[top], "bound": "hbm" if hbm_bound else "tensor", "achieved": gbs if hbm_bound else achieved,
                "peak": peak_hbm if hbm_bound else peak_tf, "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": gbs / peak_hbm if hbm_bound else achieved / peak_tf, "traffic": traffic,
                "tensor": {"achieved_tflops": achieved, "peak_tflops": peak_tf, "frac": achieved / peak_tf},
                "hbm": {"achieved_gbs": gbs, "peak_gbs": peak_hbm, "frac": gbs / peak_hbm, "algorithmic_bytes_per_step": fam_bytes[top]},
                "peak_source": "MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained (measured)" if peaks else "fallback 6.5 TB/s, 1.4 PFLOP/s",
                "launches_per_step": n_top, "algorithmic_tflop_per_step": fl_top / 1e12, "avg_launch_ms": ms_top / max(1, n_top),
                "share_of_step": ms_top / ms_step,
                "families": {names[k].split(" ")[0]: {"ms_per_step": fams[k][0], "algorithmic_tflop": fams[k][1] / 1e12, "launches": fams[k][2],
                                                      "tflops": (fams[k][1] / (fams[k][0] * 1e-3) / 1e12 if fams[k][0] > 0 else 0.0),
                                                      "gbs": (fam_bytes[k] / (fams[k][0] * 1e-3) / 1e9 if fams[k][0] > 0 else 0.0)} for k in fams},
                "note": "algorithmic FLOPs = 2*rows*out*k once per product (the 3-term TF32 split executes 3x that on the tensor pipe); "
                        "peak is the measured dense BF16 figure (TF32 is nominally half of it)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cpu_baseline

            r = cpu_baseline.iteration_sample(num_envs=N, phys_envs=256, phys_steps=2, ppo_epochs=1)
            cpu = {"value": r["env_steps_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        except Exception as ex:  # the baseline is informational; never lose the GPU number because of it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config_dict(args, world), "clocks": clock_info,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "rollout_env_steps_per_s": world * N * T / (ms_roll * 1e-3), "rollout_ms": ms_roll, "ppo_update_ms": ms_upd,
               "ppo_iteration_ms": ms_step}
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)
