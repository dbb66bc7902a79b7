#!/usr/bin/env python3
"""Per-CTA timeline of the tcgen05 GEMMs of one PPO epoch (debug build of the library, -DB200_TC_TIMELINE).

    B200_NVCC_EXTRA=-DB200_TC_TIMELINE python -m booster_gym_b200._build --force; cp booster_gym_b200/libb200t1.so scratch/lib_timeline.so
    (rebuild the product library), then on the GPU box:  python tools/tc_timeline.py scratch/lib_timeline.so

Prints, per launch (in launch order: critic fwd x3, actor fwd x3, then wgrad/dgrad alternating for actor and critic), the
median over CTAs of: total time, and where each role waited.  All times in microseconds.
"""
import ctypes as C
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    lib_dbg = sys.argv[1]
    from booster_gym_b200 import _build
    shutil.copy(lib_dbg, _build.LIB)
    os.utime(_build.LIB)
    import copy

    import torch
    import yaml

    from booster_gym_b200 import _lib
    from booster_gym_b200.learner import Learner
    from oracle import learner as L

    cfg = yaml.load(open(os.path.join(ROOT, "envs", "T1.yaml")).read(), Loader=yaml.FullLoader)
    cfg = copy.deepcopy(cfg)
    T, N = 24, int(os.environ.get("N", 4096))
    cfg["runner"]["horizon_length"] = T
    lrn = Learner(cfg, N, "cuda:0", learning_rate=1e-4, seed=1)
    lrn.load_state_dict(L.init_params(0))
    buf, last_obs, last_priv = L.synthetic_rollout(T, N, seed=3, done_rate=0.02, timeout_rate=0.03)
    dev = {k: v.cuda() for k, v in buf.items()}
    d8, t8 = dev["dones"].to(torch.uint8), dev["time_outs"].to(torch.uint8)
    lo, lp = last_obs.cuda(), last_priv.cuda()
    lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
    lib = _lib.load()
    lib.b200_tc_timeline_reset.restype = C.c_int
    lib.b200_tc_timeline_read.restype = C.c_int
    for _ in range(3):
        lrn.epoch_a(dev["rewards"], d8, t8, lo, lp)
        lrn.epoch_b(dev["actions"])
        lrn.apply()
    torch.cuda.synchronize()
    lib.b200_tc_timeline_reset()
    lrn.epoch_a(dev["rewards"], d8, t8, lo, lp)
    lrn.epoch_b(dev["actions"])
    out = np.zeros((40, 160, 12), dtype=np.uint64)
    n = lib.b200_tc_timeline_read(out.ctypes.data_as(C.c_void_p))
    names = ["c_fwd1", "c_fwd2", "c_fwd3", "a_fwd1", "a_fwd2", "a_fwd3", "a_wg3", "a_dg3", "a_wg2", "a_dg2", "a_wg1", "c_wg3", "c_dg3", "c_wg2",
             "c_dg2", "c_wg1"]
    wg = {6, 8, 10, 11, 13, 15}
    for i in range(n):
        t = out[i].astype(np.float64)
        live = t[:, 0] > 0
        t = t[live]
        t0 = t[:, 0].min()
        total = (t[:, 7].max() - t0) / 1e3
        med = lambda x: float(np.median(x)) / 1e3
        if i in wg:
            print(f"{names[i]:7s} wgrad    ctas {live.sum():3d} total {total:7.1f} | tma_done {med(t[:,1]-t[:,0]):6.1f} tma_wait_empty {med(t[:,2]):6.1f} | "
                  f"mma_done {med(t[:,3]-t[:,0]):6.1f} mma_wait_full {med(t[:,4]):6.1f} mma_wait_conv {med(t[:,5]):6.1f} | epi_start {med(t[:,6]-t[:,0]):6.1f} "
                  f"cta_end {med(t[:,7]-t[:,0]):6.1f}")
        else:
            print(f"{names[i]:7s} rowmajor ctas {live.sum():3d} total {total:7.1f} | tma_done {med(t[:,1]-t[:,0]):6.1f} tma_wait_empty {med(t[:,2]):6.1f} | "
                  f"mma_wait_tempty {med(t[:,3]):6.1f} mma_wait_full {med(t[:,4]):6.1f} mma_wait_conv {med(t[:,5]):6.1f} | epi_wait_tfull {med(t[:,6]):6.1f} "
                  f"cta_end {med(t[:,7]-t[:,0]):6.1f} | conv: wait_bempty {med(t[:,8]):5.1f} wait_afull {med(t[:,9]):5.1f} work {med(t[:,10]):5.1f} "
                  f"fence+arrive {med(t[:,11]):5.1f}")


if __name__ == "__main__":
    main()
