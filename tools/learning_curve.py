"""Learning evidence: run the unmodified training loop (`Runner.train`, the loop of utils/runner.py:99-215) with the shipped
envs/T1.yaml on one GPU and record, per iteration, what the Recorder logs: mean reward and length of the episodes that ended,
the tracking reward terms, losses, KL and learning rate.

    python tools/learning_curve.py [--iters 1000] [--num-envs 4096] [--terrain trimesh|plane] [--out profiles/r02_learning_curve.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)


def run(iters=1000, num_envs=4096, terrain=None, seed=None, quiet=True):
    import torch

    from booster_gym_b200 import _abi
    from booster_gym_b200.utils.runner import Runner

    argv = ["--task", "T1", "--num_envs", str(num_envs), "--headless", "True", "--max_iterations", str(iters)]
    if seed is not None:
        argv += ["--seed", str(seed)]
    over = {"runner": {"use_wandb": False, "save_interval": 10 ** 9}}
    if terrain:
        over["terrain"] = {"type": terrain}
    runner = Runner(test=False, argv=argv, cfg_overrides=over)
    SC = _abi.SC
    keys = ["reward", "steps", "tracking_lin_vel_x", "tracking_lin_vel_y", "tracking_ang_vel", "survival", "base_height", "orientation"]
    hist = {k: [] for k in keys + ["episodes", "kl", "lr", "value_loss", "actor_loss", "entropy"]}

    def on_iteration(it, means, count, sc):
        for k in keys:
            hist[k].append(float(means.get(k, 0.0)) if count > 0 else None)
        hist["episodes"].append(int(count))
        ep = max(1.0, sc[SC["EPOCHS"]].item())
        hist["kl"].append(sc[SC["KL"]].item())
        hist["lr"].append(sc[SC["LR"]].item())
        hist["value_loss"].append(sc[SC["SUM_VALUE_LOSS"]].item() / ep)
        hist["actor_loss"].append(sc[SC["SUM_ACTOR_LOSS"]].item() / ep)
        hist["entropy"].append(sc[SC["SUM_ENTROPY"]].item() / ep)

    t0 = time.time()
    if quiet:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            runner.train(on_iteration=on_iteration)
    else:
        runner.train(on_iteration=on_iteration)
    torch.cuda.synchronize()
    hist["wall_seconds"] = time.time() - t0
    hist["config"] = {"iterations": iters, "num_envs": num_envs, "terrain": runner.cfg["terrain"]["type"], "horizon": runner.cfg["runner"]["horizon_length"],
                      "seed": runner.cfg["basic"]["seed"]}
    return hist


def window_mean(xs, a, b):
    v = [x for x in xs[a:b] if x is not None]
    return sum(v) / len(v) if v else float("nan")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--terrain", type=str, default=None)
    ap.add_argument("--out", type=str, default=os.path.join("profiles", "r02_learning_curve.json"))
    a = ap.parse_args()
    h = run(a.iters, a.num_envs, a.terrain)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(h, open(a.out, "w"))
    n = a.iters
    print(f"wall {h['wall_seconds']:.1f} s for {n} iterations")
    for lo in range(0, n, max(1, n // 10)):
        hi = min(n, lo + max(1, n // 10))
        print(f"  it {lo:4d}-{hi:4d}: reward {window_mean(h['reward'], lo, hi):8.3f}  steps {window_mean(h['steps'], lo, hi):7.1f}  "
              f"track_x {window_mean(h['tracking_lin_vel_x'], lo, hi):7.3f}  kl {window_mean(h['kl'], lo, hi):.4f}  lr {window_mean(h['lr'], lo, hi):.2e}  "
              f"episodes/it {window_mean(h['episodes'], lo, hi):7.1f}")
