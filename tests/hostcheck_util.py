"""TEST-ONLY: build and drive tests/hostcheck/libhostcheck.so - a HOST compilation (g++ -O2 -ffp-contract=off) of the
`__host__ __device__` bodies the CUDA kernels run (csrc/t1_dynamics.cuh, t1_env.cuh, terrain.cuh, rng.cuh).  It lets the
CPU test-suite check the kernels' exact arithmetic against the golden fixtures without a GPU.  The product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
_lib = None


def lib():
    global _lib
    if _lib is None:
        csrc = os.path.join(ROOT, "booster_gym_b200", "csrc")
        deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "b200_t1.h")]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", "-o", SO, SRC])
        _lib = C.CDLL(SO)
        _lib.hc_terrain_height.restype = C.c_float
    return _lib


def pack_state(state, n):
    """reference-layout state dict -> (fstate [F_ROWS, n] float32, istate [I_ROWS, n] int32) of csrc/t1_state.h"""
    from booster_gym_b200 import _lib as L

    lib_ = L.load()
    ff, fi = L.field_table(0), L.field_table(1)
    f = np.zeros((lib_.b200_t1_num_float_rows(), n), np.float32)
    i = np.zeros((lib_.b200_t1_num_int_rows(), n), np.int32)
    for name, (row, cnt) in ff.items():
        if name in state:
            f[row:row + cnt] = np.asarray(state[name], np.float32).reshape(n, cnt).T
    for name, (row, cnt) in fi.items():
        if name in state:
            i[row:row + cnt] = np.asarray(state[name]).reshape(n, cnt).T.astype(np.int32)
    return f, i, ff, fi


def unpack(f, i, ff, fi):
    out = {}
    for name, (row, cnt) in ff.items():
        a = f[row:row + cnt].T.copy()
        out[name] = a[:, 0] if cnt == 1 else a
    for name, (row, cnt) in fi.items():
        a = i[row:row + cnt].T.copy()
        out[name] = a[:, 0] if cnt == 1 else a
    return out
