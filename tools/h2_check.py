"""h2 operand format (h2.cuh) against the 3xTF32 kernels and the fp64 oracle on the same inputs: parameter gradients of one epoch,
and the time of one epoch each way.  Usage: python tools/h2_check.py [T N] [b200_tc_set_h2 modes, e.g. 0,1,17 (17 = h2, one forward
epilogue warp group)]"""
import copy
import ctypes as C
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from booster_gym_b200 import _lib  # noqa: E402
from booster_gym_b200.learner import Learner  # noqa: E402
from oracle import learner as L  # noqa: E402


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def main():
    T, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 1000)
    modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
    cfg = copy.deepcopy(yaml.safe_load(open(os.path.join(ROOT, "envs", "T1.yaml"))))
    cfg["runner"]["horizon_length"] = T
    lib = _lib.load()
    LR = 1e-4
    lrn = Learner(cfg, N, "cuda:0", learning_rate=LR, seed=1)
    sd = L.init_params(0)
    sd["actor.6.weight"] *= 8.0
    sd["logstd"] += torch.linspace(-0.3, 0.3, 12).view(1, 12)
    lrn.load_state_dict(sd)
    buf, last_obs, last_priv = L.synthetic_rollout(T, N, seed=3, done_rate=0.02, timeout_rate=0.03)
    with torch.no_grad():
        mu0 = L.actor_mean(sd, buf["obses"])
        buf["actions"] = mu0 + torch.exp(sd["logstd"]) * torch.randn(T, N, 12, generator=torch.Generator().manual_seed(5))
    dev = {k: v.cuda() for k, v in buf.items()}
    d8, t8 = dev["dones"].to(torch.uint8), dev["time_outs"].to(torch.uint8)
    lo, lp = last_obs.cuda(), last_priv.cuda()
    # fp64 oracle, one epoch
    sdd = {k: v.double().clone() for k, v in sd.items()}
    bufd = {k: (v.double().clone() if v.is_floating_point() else v.clone()) for k, v in buf.items()}
    omu, osig, olp = L.old_dist(sdd, bufd["obses"], bufd["actions"])
    o64 = L.epoch(sdd, L.new_adam(sdd), bufd, last_obs.double(), last_priv.double(), omu, osig, olp, LR)
    snaps = {}
    for m in modes:
        lib.b200_tc_set_h2(m)
        lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
        rew = dev["rewards"].clone()
        lrn.epoch_a(rew, d8, t8, lo, lp)
        lrn.epoch_b(dev["actions"])
        torch.cuda.synchronize()
        snaps[m] = ({k: v.clone().cpu() for k, v in lrn.views(lrn.grads).items()}, lrn.buffer(3, (T * N, 12)).clone().cpu(), lrn.buffer(0, (T * N,)).clone().cpu())
        print(f"h2 mode {m}: epoch ran; mu vs fp64 {rel(snaps[m][1], o64['mu'].reshape(T * N, 12)):.3e}  V vs fp64 {rel(snaps[m][2], o64['values'].reshape(T * N)):.3e}", flush=True)
    for name in o64["grads"]:
        line = f"  grad {name:18s}"
        for m in modes:
            line += f"  mode {m} vs fp64 {rel(snaps[m][0][name].reshape(o64['grads'][name].shape), o64['grads'][name]):.3e}"
        print(line)
    for m in modes:
        lib.b200_tc_set_h2(m)
        for _ in range(2):
            lrn.epoch_a(rew, d8, t8, lo, lp)
            lrn.epoch_b(dev["actions"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lrn.epoch_a(rew, d8, t8, lo, lp)
            lrn.epoch_b(dev["actions"])
        e1.record()
        torch.cuda.synchronize()
        print(f"h2 mode {m}: {e0.elapsed_time(e1) / 5 * 1000:.1f} us per epoch (T={T}, N={N})")
        lib.b200_profile_gemm(1)
        lrn.epoch_a(rew, d8, t8, lo, lp)
        lrn.epoch_b(dev["actions"])
        torch.cuda.synchronize()
        for kind, nm in ((2, "wgrad"), (3, "k_mlp_fwd"), (4, "k_mlp_bwd")):
            ms, fl, n = C.c_double(), C.c_double(), C.c_int()
            lib.b200_profile_gemm_read(kind, C.byref(ms), C.byref(fl), C.byref(n))
            if n.value:
                print(f"    {nm:14s} {n.value:3d} launches {ms.value * 1000:8.1f} us  {fl.value / max(ms.value, 1e-9) / 1e9:7.1f} TFLOP/s algorithmic")
        lib.b200_profile_gemm(0)
    lib.b200_tc_set_h2(0)


if __name__ == "__main__":
    main()
