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
`--impl reference` times the CPU port of the same iteration (oracle/cpu_baseline.py) on the host cores: whole iterations at a
bounded number of environments, measured as they run.
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
            "l2": "working set (the learner streams ~2.2 GB of activations and gradients per epoch through a >1 GB workspace per GPU) is larger than "
                  "the 126 MB L2; no flush needed",
            "rollout_cuda_graph": bool(args.graphs),
            "update_cuda_graph": bool(args.graphs) and os.environ.get("B200_UPDATE_GRAPH", "1") != "0"
                                 and (world == 1 or os.environ.get("B200_PEER_EXCHANGE", "1") != "0")}


# ---------------------------------------------------------------------------------------------------------------------
REF_ENVS = 1024   # environments of the CPU arms' bounded sample (a whole iteration at a quarter of configs[1]'s 4096 envs)


def reference_arm(args):
    """The path's CPU implementation on the host cores (oracle/cpu_baseline.py: the port - neither reference physics engine exists
    on the box); rank 0 only.  Each step is one WHOLE iteration (24-step rollout incl. policy, physics, post-physics pass +
    old-dist + 20 epochs) at REF_ENVS environments, timed as it runs: ms_per_step x steps is wall time spent inside this process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline

    it = cpu_baseline.CpuIteration(num_envs=REF_ENVS)
    for _ in range(args.warmup):
        it.iteration()
    secs, parts = [], None
    for _ in range(max(1, args.steps)):
        s_, parts = it.iteration()
        secs.append(s_)
    sec = sum(secs) / len(secs)
    value = REF_ENVS * 24 / sec
    cfg = config_dict(args, 1)
    cfg["reference_sample_envs"] = REF_ENVS
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(secs), "warmup": args.warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "reference", "config": cfg,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": it.cores, "kind": "port", "sample": it.sample_text()},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "one CPU process on rank 0 regardless of --gpus; every step is a whole iteration at the sample's env count, "
                   "measured, not extrapolated (CPU cost is linear in the number of environments)",
           "parts_s": parts}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
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

    # the update as a second graph (not with the NCCL protocol, whose collectives are issued from Python between the kernels); lr / Adam
    # step / KL rule / peer-exchange sequence numbers are device-resident, so a replay IS the next update
    upd_graph, upd_launches = None, 0
    if args.graphs and (world == 1 or lrn.peers_bound) and os.environ.get("B200_UPDATE_GRAPH", "1") != "0":
        runner.update(obs, priv)
        torch.cuda.synchronize(dev)
        upd_graph = torch.cuda.CUDAGraph()
        lc0 = lib.b200_launch_count()
        with torch.cuda.graph(upd_graph):
            runner.update(obs, priv)
        upd_launches = lib.b200_launch_count() - lc0

    def update():
        if upd_graph is not None:
            upd_graph.replay()
        else:
            runner.update(obs, priv)

    def iteration():
        if graph is not None:
            graph.replay()
        else:
            runner.rollout(obs, priv)
        update()

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

    # nvidia-smi needs a few hundred ms to start reporting (longer on an 8-GPU box) and the default timed region lasts ~0.1 s:
    # the sampler runs from before the warm-up to after the timed region, and the same iteration keeps the GPU under the same load
    # (untimed) until it has seen it for at least a second
    clocks = ClockSampler(local)
    t_load = time.perf_counter()
    for _ in range(max(3, args.warmup)):
        iteration()
    # ---- main measurement: states resident in HBM --------------------------------------------------------------
    l0 = lib.b200_launch_count()
    ms_total = timed(iteration, args.steps)
    launches = lib.b200_launch_count() - l0 + (graph_launches * args.steps if graph is not None else 0) + upd_launches * args.steps
    while time.perf_counter() - t_load < 1.2:
        iteration()
    torch.cuda.synchronize(dev)
    clock_info = clocks.stop()
    ms_step = ms_total / args.steps
    value = world * N * T / (ms_step * 1e-3)

    # ---- phase split (rollout only / update only), informational ----------------------------------------------------
    ms_roll = timed(lambda: (graph.replay() if graph is not None else runner.rollout(obs, priv)), max(2, args.steps)) / max(2, args.steps)
    ms_upd = timed(update, max(2, args.steps)) / max(2, args.steps)

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

    # ---- roofline of the dominant kernel family: CUDA events around every launch of the MLP GEMM kernels (recorded inside the
    # library on the launching stream) during one extra iteration.  SURVEY 8(d) assigns K6 (the MLP contractions) to the TENSOR
    # roof, so that is `bound`; the HBM side is reported beside it.  Algorithmic figures per sample per epoch from SURVEY 8(d):
    # 1 005 312 FLOP (fwd 353 536 + bwd 651 776) and 304 B if activations stayed on chip; `traffic_model` is what this
    # implementation's launches are DESIGNED to move (every operand once, every result once), `traffic` what ncu measured.
    lib.b200_profile_gemm(1)
    if graph is not None:
        graph.replay()
    else:
        runner.rollout(obs, priv)
    runner.update(obs, priv)     # eager: the events are recorded by the library around its launches
    torch.cuda.synchronize(dev)
    h2 = os.environ.get("B200_H2", "1") != "0" and os.environ.get("B200_CHAIN", "1") == "1"
    if h2:   # default: the h2 operand format (two fp16 halves per word, tcgen05.mma kind::f16; h2.cuh, mlp_chain_h2.cuh)
        names = {2: ("k_wgrad_h2", "all six MLP weight gradients in ONE launch: tcgen05 kind::f16 on MN-major h2 words straight from TMA, register flush every 64 MMAs"),
                 3: ("k_mlp_fwd_h2", "fused forward layer chain on h2 words: 3 hidden layers of both nets per launch, hand-over = one packed word per TMEM column (TS-form tcgen05.mma kind::f16)"),
                 4: ("k_mlp_bwd_h2", "fused backward layer chain on h2 words: dz3 -> dz2 -> dz1 + bias gradients of both nets per launch")}
        split_exec, split_name = 4.0, "h2 split (2 MMAs of twice the K per product = 4x the algorithmic FLOPs at the BF16 / FP16 rate)"
    else:
        names = {1: ("k_tc_rowmajor", "tcgen05 3xTF32 GEMM, one launch per layer (B200_CHAIN=0 path)"),
                 2: ("k_tc_wgrad", "tcgen05/TMEM/TMA 3xTF32, MN-major operands, in-smem split: the six MLP weight gradients"),
                 3: ("k_mlp_fwd", "fused forward layer chain: 3 hidden layers per launch, activations handed over in TMEM (TS-form tcgen05.mma)"),
                 4: ("k_mlp_bwd", "fused backward layer chain: dz3 -> dz2 -> dz1 + bias gradients per launch")}
        split_exec, split_name = 6.0, "3xTF32 split (3 MMAs per product at the TF32 rate = 6x the algorithmic FLOPs at the BF16 rate)"
    fams = {}
    for kind in names:
        ms_g, fl_g, n_g, by_g = C.c_double(), C.c_double(), C.c_int(), C.c_double()
        lib.b200_profile_gemm_read(kind, C.byref(ms_g), C.byref(fl_g), C.byref(n_g))
        lib.b200_profile_gemm_bytes(kind, C.byref(by_g))
        if n_g.value:
            fams[kind] = {"ms": ms_g.value, "flop": fl_g.value, "launches": n_g.value, "bytes_model": by_g.value}
    lib.b200_profile_gemm(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_hbm = peaks.get("hbm_gbs", 6500.0)
    ncu_traffic = {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this round
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        pass
    top = max(fams, key=lambda k: fams[k]["ms"])
    M = T * N
    fam_out = {}
    for k, f in fams.items():
        tfl = f["flop"] / (f["ms"] * 1e-3) / 1e12
        fam_out[names[k][0]] = {"ms_per_step": f["ms"], "launches": f["launches"], "algorithmic_tflop": f["flop"] / 1e12, "tflops": tfl,
                                "frac_of_bf16_sustained": tfl / peak_tf, "frac_of_split_ceiling": tfl / (peak_tf / split_exec),
                                "traffic_model_gbs": f["bytes_model"] / (f["ms"] * 1e-3) / 1e9,
                                "traffic_model_frac_of_hbm_peak": f["bytes_model"] / (f["ms"] * 1e-3) / 1e9 / peak_hbm,
                                "traffic_model_bytes_per_launch": f["bytes_model"] / f["launches"],
                                "ncu_dram_bytes_per_launch": ncu_traffic.get(names[k][0])}
    ft = fams[top]
    achieved = ft["flop"] / (ft["ms"] * 1e-3) / 1e12
    all_ms = sum(f["ms"] for f in fams.values())
    all_fl = sum(f["flop"] for f in fams.values())
    roofline = {"kernel": f"{names[top][0]} ({names[top][1]})", "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": ncu_traffic.get(names[top][0]),
                "frac_of_split_ceiling": achieved / (peak_tf / split_exec), "split_ceiling_tflops": peak_tf / split_exec,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs (measured on this pool)" if peaks
                                else "fallback 1.4 PFLOP/s, 6.5 TB/s (B200_PROFILING.md)"),
                "launches_per_step": ft["launches"], "algorithmic_flop_per_launch": ft["flop"] / ft["launches"],
                "avg_launch_ms": ft["ms"] / ft["launches"], "share_of_step": ft["ms"] / ms_step,
                "hbm": {"algorithmic_bytes_per_sample_epoch": 304, "algorithmic_bytes_per_step": 304.0 * M * E,
                        "traffic_model_bytes_per_step": sum(f["bytes_model"] for f in fams.values()),
                        "traffic_model_gbs_all_mlp_kernels": sum(f["bytes_model"] for f in fams.values()) / (all_ms * 1e-3) / 1e9,
                        "peak_gbs": peak_hbm},
                "all_mlp_kernels": {"ms_per_step": all_ms, "algorithmic_tflop_per_step": all_fl / 1e12, "tflops": all_fl / (all_ms * 1e-3) / 1e12,
                                    "frac_of_bf16_sustained": all_fl / (all_ms * 1e-3) / 1e12 / peak_tf,
                                    "share_of_step": all_ms / ms_step},
                "families": fam_out,
                "note": "algorithmic FLOPs = 2*rows*out*k once per product; the fp32-accurate " + split_name + " bounds this arithmetic class at "
                        "1/%d of the measured BF16 peak (frac_of_split_ceiling); `traffic` = dram bytes per launch of this round's ncu capture" % int(split_exec)}

    # ---- k_physics against the MEASURED FP32 FMA peak (FP32-pipe roofline, measured on the box)
    fma = C.c_double()
    lib.b200_fma_peak(C.byref(fma), stream.cuda_stream)
    FLOP_TICK = 13218.0   # tools/count_flops.py (instrumented exact count, standing robot; frozen in BASELINE.md section 5)
    act0 = torch.zeros(N, 12, device=dev)
    for _ in range(3):
        env.physics(act0, 10, apply_pd=True)
    torch.cuda.synchronize(dev)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record(stream)
    for _ in range(20):
        env.physics(act0, 10, apply_pd=True)
    pe1.record(stream)
    torch.cuda.synchronize(dev)
    phys_ms = pe0.elapsed_time(pe1) / 20
    phys_tf = FLOP_TICK * 10 * N / (phys_ms * 1e-3) / 1e12
    roofline["k_physics"] = {"bound": "fp32", "avg_launch_ms": phys_ms, "flop_per_env_tick": FLOP_TICK, "achieved_tflops": phys_tf,
                             "peak_tflops_measured_fma": fma.value, "frac": phys_tf / fma.value if fma.value > 0 else None,
                             "note": "one launch = 10 ticks of every env; latency-bound at this env count (2 lanes per env, 1 warp per CTA)"}
    exchange = "none" if world == 1 else ("peer" if lrn.peers_bound else "nccl")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cpu_baseline

            r = cpu_baseline.iteration_sample(num_envs=REF_ENVS, iterations=1, warmup=1)
            cpu = {"value": r["env_steps_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "seconds_per_iteration_at_sample_size": r["seconds_per_iteration"], "parts_s": r["parts"]}
        except Exception as ex:  # the baseline is informational; never lose the GPU number because of it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config_dict(args, world), "clocks": clock_info,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "rollout_env_steps_per_s": world * N * T / (ms_roll * 1e-3), "rollout_ms": ms_roll, "ppo_update_ms": ms_upd,
               "ppo_iteration_ms": ms_step, "exchange": exchange,
               "mlp_path": "h2" if h2 else ("chain-tf32" if os.environ.get("B200_CHAIN", "1") != "0" else "layers-tf32")}
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)
