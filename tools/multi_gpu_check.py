#!/usr/bin/env python3
"""Multi-GPU parity of the learner's two exchange protocols (run under torchrun, >= 2 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py

Each rank owns N envs of a synthetic rollout.  Two learners with identical parameters run 3 epochs: (A) the NCCL protocol of
Runner.update (all-reduce of the advantage moments, the gradient and the loss sums), (B) the NVLink peer-memory exchange
inside the library (b200_ppo_bind_peers).  Rank 0 additionally holds the fp64 oracle's full-batch result (all ranks' samples
in one process).  Checks: B == A to fp32 summation-order noise (parameters after 3 Adam steps, learning rate, KL), every
rank holds identical parameters, and both match the full-batch oracle within the single-GPU tolerances.
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import yaml

    from booster_gym_b200 import _abi
    from booster_gym_b200.learner import Learner
    from oracle import learner as L

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    cfg = yaml.load(open(os.path.join(ROOT, "envs", "T1.yaml")).read(), Loader=yaml.FullLoader)
    cfg = copy.deepcopy(cfg)
    T, N = 6, 512
    cfg["runner"]["horizon_length"] = T
    LR = 1e-4
    sd = L.init_params(0)
    sd["actor.6.weight"] *= 8.0
    buf, last_obs, last_priv = L.synthetic_rollout(T, N * world, seed=3, done_rate=0.02, timeout_rate=0.03)
    with torch.no_grad():
        buf["actions"] = L.actor_mean(sd, buf["obses"]) + torch.exp(sd["logstd"]) * torch.randn(T, N * world, 12, generator=torch.Generator().manual_seed(5))
    sl = slice(rank * N, (rank + 1) * N)
    mine = {k: v[:, sl].contiguous().to(dev) for k, v in buf.items()}
    lo, lp = last_obs[sl].contiguous().to(dev), last_priv[sl].contiguous().to(dev)
    d8, t8 = mine["dones"].to(torch.uint8), mine["time_outs"].to(torch.uint8)

    def run(peer, graphed=False):
        lrn = Learner(cfg, N, dev, world_size=world, env_base=rank * N, learning_rate=LR, seed=1)
        lrn.load_state_dict(sd)
        if peer:
            assert lrn.bind_peers(), "peer-memory exchange could not be bound"
        rew = mine["rewards"].clone()
        lrn.old_dist(mine["obses"], mine["privileged_obses"], mine["actions"])
        if graphed:
            # (C) one epoch of the peer protocol captured in a CUDA graph and replayed three times (Runner.update_graphed / bench.py under
            # torchrun): the exchange sequence numbers and slot parities advance on the device
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                lrn.epoch_a(rew, d8, t8, lo, lp)
                lrn.epoch_b(mine["actions"])
                lrn.apply()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize(dev)
            lrn._graph = g
            return lrn
        for ep in range(3):
            if ep == 1:
                lrn.after_first_epoch = {k: v.clone() for k, v in lrn.views().items()}
            lrn.epoch_a(rew, d8, t8, lo, lp)
            if not peer:
                dist.all_reduce(lrn.dstats[0:4])
            lrn.epoch_b(mine["actions"])
            if not peer:
                dist.all_reduce(lrn.grads)
                dist.all_reduce(lrn.dstats[4:10])
            lrn.apply()
        torch.cuda.synchronize(dev)
        return lrn

    a, b = run(False), run(True)
    c = run(True, graphed=True)
    worst_c = 0.0
    for name in b.views():
        pb, pc = b.views()[name], c.views()[name]
        worst_c = max(worst_c, (pb - pc).abs().max().item() / max(pb.abs().max().item(), 1e-12))
    assert worst_c <= 2e-6, ("graph replay of the peer protocol", worst_c)
    assert c.scalars[_abi.SC["ADAM_STEP"]].item() == 3
    worst = 0.0
    for name in a.views():
        pa, pb = a.views()[name], b.views()[name]
        worst = max(worst, (pa - pb).abs().max().item() / max(pa.abs().max().item(), 1e-12))
    sc_a, sc_b = a.scalars.cpu(), b.scalars.cpu()
    for key in ("LR", "KL", "VALUE_LOSS", "ACTOR_LOSS", "GRAD_NORM", "ADV_MEAN", "ADV_STD"):
        va, vb = sc_a[_abi.SC[key]].item(), sc_b[_abi.SC[key]].item()
        assert abs(va - vb) <= 1e-5 * max(abs(va), 1e-3), (key, va, vb)
    assert worst <= 2e-6, worst            # 3 Adam steps of 1e-4 on fp32 sums taken in a different order
    # every rank holds the same parameters
    flat = b.params.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref), "ranks diverged"
    # full-batch fp64 oracle (all ranks' samples in one process)
    if rank == 0:
        sdd = {k: v.double().clone() for k, v in sd.items()}
        bufd = {k: (v.double().clone() if v.is_floating_point() else v.clone()) for k, v in buf.items()}
        omu, osig, olp = L.old_dist(sdd, bufd["obses"], bufd["actions"])
        adam, lr = L.new_adam(sdd), LR
        outs, lrs, sd_after_first = [], [], None
        for ep in range(3):
            if ep == 1:
                sd_after_first = {k: v.clone() for k, v in sdd.items()}
            lrs.append(lr)
            o = L.epoch(sdd, adam, bufd, last_obs.double(), last_priv.double(), omu, osig, olp, lr)
            lr = o["lr"]
            outs.append(o)
        # Full-batch equivalence is a statement about ONE epoch (sharded moments / gradients / loss sums == the full batch's): the
        # parameters after the first Adam step are held to the single-GPU element-wise bound.  Adam's update lr * m_hat / (sqrt(v_hat)
        # + eps) is ~ lr * sign(g), so an element whose gradient is small against the stated gradient tolerance (2e-4 of the tensor
        # max, tests/test_gpu_learner.py) may legitimately move differently: lr * min(2, 2 * 2e-4 * max|g| / |g_ij|), on top of 1e-5
        # relative.  Epochs 2 and 3 start from parameters that already differ by that much and this synthetic batch drives KL to 0.28
        # (far outside the trust region), so their errors compound: after three steps lr, KL and the loss scalars must still agree
        # and the parameters stay within the three steps' total movement.
        werr1, wviol = 0.0, 0.0
        for name, ref64 in sd_after_first.items():
            ours = b.after_first_epoch[name].cpu().double().reshape(ref64.shape)
            err = (ours - ref64).abs()
            g = outs[0]["grads"][name].double().reshape(err.shape).abs()
            bound = lrs[0] * torch.clamp(2.0 * 2e-4 * g.max() / g.clamp_min(1e-30), max=2.0)
            werr1 = max(werr1, err.max().item())
            wviol = max(wviol, (err - bound - 1e-5 * ref64.abs().max()).max().item())
        assert wviol <= 0.0, (wviol, werr1)
        werr = max((b.views()[name].cpu().double().reshape(ref64.shape) - ref64).abs().max().item() for name, ref64 in sdd.items())
        assert werr <= sum(lrs), (werr, lrs)
        assert abs(sc_b[_abi.SC["LR"]].item() - lr) <= 1e-6 * lr
        # (KL after THREE compounding epochs: measured 1.2e-4 relative at world size 8, 2e-5 at 2 and 4)
        assert abs(sc_b[_abi.SC["KL"]].item() - o["kl"]) <= 1e-3 * max(abs(o["kl"]), 1e-3) + 1e-7, (sc_b[_abi.SC["KL"]].item(), o["kl"])
        print(f"MULTI-GPU OK world={world}: peer vs NCCL {worst:.2e}, graph-replayed peer vs eager peer {worst_c:.2e}, vs fp64 full batch {werr1:.2e} after one epoch / {werr:.2e} after three (lr {lr:.3e})")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
