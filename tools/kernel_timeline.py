"""In-situ kernel durations of a few whole PPO iterations (rollout graph + 20-epoch update) from CUPTI activity records
(torch.profiler sees every kernel of the process, also those libb200t1.so launches): unlike an ncu launch list the kernels
run back to back with warm caches, so the per-kernel AVERAGES are the ones that add up to the iteration time.

    python tools/kernel_timeline.py [--num-envs 4096] [--terrain plane] [--iters 3] [--out gpurun_out/timeline.json]
"""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--terrain", type=str, default="plane")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--out", type=str, default=None)
    ap.add_argument("--gaps", action="store_true", help="list the largest idle gaps between consecutive kernels")
    ap.add_argument("--eager-update", action="store_true", help="launch the update kernel by kernel instead of replaying its CUDA graph")
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile

    from booster_gym_b200.utils.runner import Runner

    argv = ["--task", "T1", "--num_envs", str(a.num_envs), "--headless", "True", "--max_iterations", "1"]
    runner = Runner(test=False, argv=argv, cfg_overrides={"runner": {"use_wandb": False, "save_interval": 10 ** 9}, "terrain": {"type": a.terrain}})
    obs, infos = runner.env.reset()
    priv = infos["privileged_obs"]

    def iteration(obs, priv):
        obs, priv = runner.rollout_graphed(obs, priv)
        (runner.update if a.eager_update else runner.update_graphed)(obs, priv)
        return obs, priv

    for _ in range(4):
        obs, priv = iteration(obs, priv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        obs, priv = iteration(obs, priv)
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / a.iters
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.iters):
            obs, priv = iteration(obs, priv)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    t_first, t_last = None, None
    for ev in prof.events():
        if ev.device_type.name != "CUDA" if hasattr(ev.device_type, "name") else False:
            continue
        dur = getattr(ev, "device_time", 0) or getattr(ev, "cuda_time", 0)
        if not dur:
            continue
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "")
        name = re.sub(r"^b200::", "", name)
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += dur
    if a.gaps:
        # idle time between consecutive kernels on the device, by (previous kernel -> next kernel)
        evs = []
        for ev in prof.events():
            dur = getattr(ev, "device_time", 0) or getattr(ev, "cuda_time", 0)
            if dur and hasattr(ev, "time_range"):
                evs.append((ev.time_range.start, ev.time_range.end, re.sub(r"\(.*", "", ev.name).replace("void ", "")))
        evs.sort()
        gaps = collections.OrderedDict()
        end_prev, name_prev = None, None
        for st, en, nm in evs:
            if end_prev is not None and st > end_prev:
                g = gaps.setdefault((name_prev[:36], nm[:36]), [0, 0.0])
                g[0] += 1
                g[1] += st - end_prev
            if end_prev is None or en > end_prev:
                end_prev, name_prev = en, nm
        print(f"idle gaps per iteration: {sum(v[1] for v in gaps.values()) / a.iters / 1e3:.3f} ms")
        for (p_, n_), (cnt, us) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f"   {p_:36s} -> {n_:36s} {cnt / a.iters:6.1f} per it  {us / cnt:8.2f} us avg  {us / a.iters:9.1f} us/it")
    tot = sum(v[1] for v in agg.values())
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    print(f"iteration (CUDA events, no profiler): {ms_plain:.3f} ms; sum of kernel durations under CUPTI: {tot / a.iters / 1e3:.3f} ms per iteration")
    print(f"{'kernel':60s} {'per it':>7s} {'avg us':>9s} {'us/it':>9s} {'share':>6s}")
    for k, (n, us) in rows[:40]:
        print(f"{k[:60]:60s} {n / a.iters:7.1f} {us / n:9.2f} {us / a.iters:9.1f} {us / tot:6.3f}")
    if a.out:
        json.dump({"ms_per_iteration": ms_plain, "kernel_ms_per_iteration": tot / a.iters / 1e3, "num_envs": a.num_envs, "terrain": a.terrain,
                   "kernels": {k: {"per_iteration": n / a.iters, "avg_us": us / n, "us_per_iteration": us / a.iters} for k, (n, us) in rows}},
                  open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
