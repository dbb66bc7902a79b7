"""Fused layer chains (mlp_chain.cuh) against the layer-by-layer tcgen05 path on the same inputs: activations, input
gradients, parameter gradients, and the time of one epoch each way.  Usage: python tools/chain_check.py [T N]"""
import copy
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
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def main():
    T, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 1000)
    modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]   # 0 layers, 1 chain, 2 chain on CTA pairs
    cfg = yaml.safe_load(open(os.path.join(ROOT, "envs", "T1.yaml")))
    cfg = copy.deepcopy(cfg)
    cfg["runner"]["horizon_length"] = T
    lib = _lib.load()
    lrn = Learner(cfg, N, "cuda:0", learning_rate=1e-4, seed=1)
    sd = L.init_params(0)
    sd["actor.6.weight"] *= 8.0
    lrn.load_state_dict(sd)
    buf, last_obs, last_priv = L.synthetic_rollout(T, N, seed=3, done_rate=0.02, timeout_rate=0.03)
    with torch.no_grad():
        mu0 = L.actor_mean(sd, buf["obses"])
        buf["actions"] = mu0 + torch.exp(sd["logstd"]) * torch.randn(T, N, 12, generator=torch.Generator().manual_seed(5))
    dev = {k: v.cuda() for k, v in buf.items()}
    d8, t8 = dev["dones"].to(torch.uint8), dev["time_outs"].to(torch.uint8)
    lo, lp = last_obs.cuda(), last_priv.cuda()
    M = T * N
    shapes = {0: (M,), 3: (M, 12), 7: (M + N, 256), 8: (M + N, 256), 9: (M + N, 128), 10: (M, 256), 11: (M, 128), 12: (M, 128),
              13: (M, 256), 14: (M, 256), 5: (M,), 6: (M, 12)}
    names = {0: "V", 3: "mu", 7: "C1", 8: "C2", 9: "C3", 10: "A1", 11: "A2", 12: "A3", 13: "dz2 critic", 14: "dz1 critic", 5: "dV", 6: "dmu"}
    snaps = {}
    for chain in modes:
        lib.b200_tc_set_chain(chain)
        lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
        rew = dev["rewards"].clone()
        lrn.epoch_a(rew, d8, t8, lo, lp)
        lrn.epoch_b(dev["actions"])
        torch.cuda.synchronize()
        snaps[chain] = ({k: lrn.buffer(k, s) for k, s in shapes.items()}, {k: v.clone() for k, v in lrn.views(lrn.grads).items()})
        print(f"chain={chain}: epoch ran", flush=True)
    base = modes[0]
    for m in modes[1:]:
        for k in shapes:
            print(f"  {names[k]:12s} mode {m} vs mode {base}: {rel(snaps[m][0][k], snaps[base][0][k]):.3e}")
        for k in snaps[base][1]:
            print(f"  grad {k:18s} mode {m} vs mode {base}: {rel(snaps[m][1][k], snaps[base][1][k]):.3e}")
    # the chain's input gradients against torch fp64 on the chain's own inputs; where are the bad rows?
    lib.b200_tc_set_chain(modes[-1])
    rew = dev["rewards"].clone()
    lrn.epoch_a(rew, d8, t8, lo, lp)
    lrn.epoch_b(dev["actions"])
    torch.cuda.synchronize()
    P = lrn.views()
    for net, (iz3, ih2, ih1, iz2, iz1, n2, w3, w2) in {"critic": (17, 8, 7, 13, 14, 256, "critic.4.weight", "critic.2.weight"),
                                                         "actor": (18, 11, 10, 15, 16, 128, "actor.4.weight", "actor.2.weight")}.items():
        z3 = lrn.buffer(iz3, (M, 128)).double()
        h2 = lrn.buffer(ih2, (M + N if net == "critic" else M, n2))[:M].double()
        h1 = lrn.buffer(ih1, (M + N if net == "critic" else M, 256))[:M].double()
        z2 = lrn.buffer(iz2, (M, n2)).double()
        z1 = lrn.buffer(iz1, (M, 256)).double()
        e2 = (z3 @ P[w3].double()) * torch.where(h2 > 0, torch.ones_like(h2), h2 + 1)
        e1 = (e2 @ P[w2].double()) * torch.where(h1 > 0, torch.ones_like(h1), h1 + 1)
        for nm, got, exp in (("dz2", z2, e2), ("dz1", z1, e1)):
            err = (got - exp).abs() / exp.abs().max()
            bad = (err > 1e-4).any(dim=1).nonzero().flatten()
            print(f"  {net} {nm}: max err {err.max().item():.3e}; bad rows {bad.numel()} / {M}; bad tiles {sorted(set((bad // 128).tolist()))[:40]}")
            if bad.numel():
                r = bad[0].item()
                cols = (err[r] > 1e-4).nonzero().flatten().tolist()
                print(f"     first bad row {r} (tile {r // 128}, row in tile {r % 128}) bad cols {cols[:24]} ... n={len(cols)}")
    # fp64 oracle for the activations of the critic / actor (the absolute yardstick)
    sd64 = {k: v.double() for k, v in sd.items()}
    mu64 = L.actor_mean(sd64, buf["obses"].double()).reshape(M, 12)
    v64 = L.critic_value(sd64, buf["obses"].double(), buf["privileged_obses"].double()).reshape(M)
    for chain in modes:
        print(f"  chain={chain}: mu vs fp64 {rel(snaps[chain][0][3].cpu().double(), mu64):.3e}   V vs fp64 {rel(snaps[chain][0][0].cpu().double(), v64):.3e}")
    for chain in modes:
        lib.b200_tc_set_chain(chain)
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
        print(f"chain={chain}: {e0.elapsed_time(e1) / 5 * 1000:.1f} us per epoch (T={T}, N={N})")
        import ctypes as C
        lib.b200_profile_gemm(1)
        lrn.epoch_a(rew, d8, t8, lo, lp)
        lrn.epoch_b(dev["actions"])
        torch.cuda.synchronize()
        for kind, nm in ((1, "k_tc_rowmajor"), (2, "k_tc_wgrad"), (3, "k_mlp_fwd"), (4, "k_mlp_bwd")):
            ms, fl, n = C.c_double(), C.c_double(), C.c_int()
            lib.b200_profile_gemm_read(kind, C.byref(ms), C.byref(fl), C.byref(n))
            if n.value:
                print(f"    {nm:14s} {n.value:3d} launches {ms.value * 1000:8.1f} us  {fl.value / max(ms.value, 1e-9) / 1e9:7.1f} TFLOP/s algorithmic")
        lib.b200_profile_gemm(0)
    lib.b200_tc_set_chain(1)


if __name__ == "__main__":
    main()
