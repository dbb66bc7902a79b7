"""Where the fused-chain kernels wait (debug build: B200_NVCC_EXTRA=-DB200_CHAIN_TL python -m booster_gym_b200._build --force).
Prints, averaged over CTAs, the SM-clock cycles the MMA-issuing thread and the first epilogue warp spent blocked per barrier
class during ONE launch of k_mlp_fwd (critic + actor when --both) and k_mlp_bwd at T x N samples."""
import copy
import ctypes as C
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from booster_gym_b200 import _lib  # noqa: E402
from booster_gym_b200.learner import Learner  # noqa: E402
from oracle import learner as L  # noqa: E402

T, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24, 4096)
cfg = copy.deepcopy(yaml.safe_load(open(os.path.join(ROOT, "envs", "T1.yaml"))))
cfg["runner"]["horizon_length"] = T
lib = _lib.load()
lib.b200_chain_timeline_read.restype = C.c_int
lib.b200_chain_timeline_read.argtypes = [C.c_void_p]
if len(sys.argv) > 3:
    lib.b200_tc_set_h2(int(sys.argv[3]))   # e.g. 35 = h2 chains, two epilogue warp groups
lrn = Learner(cfg, N, "cuda:0", learning_rate=1e-4, seed=1)
lrn.load_state_dict(L.init_params(0))
buf, lo, lp = L.synthetic_rollout(T, N, seed=3, done_rate=0.02, timeout_rate=0.03)
dev = {k: v.cuda() for k, v in buf.items()}
d8, t8 = dev["dones"].to(torch.uint8), dev["time_outs"].to(torch.uint8)
lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
for _ in range(3):
    lrn.epoch_a(dev["rewards"], d8, t8, lo.cuda(), lp.cuda())
    lrn.epoch_b(dev["actions"])
out = np.zeros((2, 160, 16), dtype=np.int64)
lib.b200_chain_timeline_read(out.ctypes.data)
for k, name in ((0, "k_mlp_fwd (last launch = actor)"), (1, "k_mlp_bwd (critic + actor)")):
    a = out[k][:148].astype(np.float64)
    a = a[a[:, 5] > 0]
    print(name, f"CTAs {len(a)}")
    m = a.mean(0)
    print(f"  MMA thread: total {m[5]:9.0f} clk | wait x_full {m[0]:8.0f}  accf-safety {m[1]:8.0f}  weights {m[2]:8.0f}  lo_full(L2/dgrad3) {m[3]:8.0f}  lo_full(L3/dgrad2) {m[4]:8.0f}"
          f"  -> issuing/other {m[5] - m[:5].sum():8.0f}")
    if k == 0:
        print(f"  epilogue  : total {m[15]:9.0f} clk | wait accf0 {m[8]:8.0f}  accf1 {m[9]:8.0f}  accf2 {m[10]:8.0f}  lo_empty {m[11]:8.0f} | st.global {m[12]:8.0f}  "
              f"publish_slice (split, tcgen05.st, st.shared) {m[13]:8.0f}  publish_done (wait::st, fences, arrive) {m[14]:8.0f}  -> tcgen05.ld + bias + ELU {m[15] - m[8:15].sum():8.0f}")
    else:
        print(f"  epilogue  : total {m[12]:9.0f} clk | wait accf0 {m[8]:8.0f}  accf1 {m[9]:8.0f}  aux {m[10]:8.0f}  lo_empty {m[11]:8.0f}  -> working {m[12] - m[8:12].sum():8.0f}")
    print(f"  total max {a[:, 5].max():.0f} min {a[:, 5].min():.0f}")
