"""ctypes front-end of oracle/physics_oracle.c (FP64 CPU oracle).  TEST INFRASTRUCTURE ONLY - see the C header."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "physics_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off", "-o", _SO, src, "-lm"])
    return _SO


_PORT_SO = os.path.join(_HERE, "liboracle_port.so")


def build_port(force=False):
    """the -O3 host build of the product's recursion used ONLY as the timed CPU baseline (oracle/physics_port.cpp)"""
    src = os.path.join(_HERE, "physics_port.cpp")
    deps = [src, os.path.join(_HERE, "..", "booster_gym_b200", "csrc", "t1_dynamics.cuh")]
    if force or not os.path.exists(_PORT_SO) or any(os.path.getmtime(_PORT_SO) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", _PORT_SO, src, "-lm"])
    return _PORT_SO


_port = None


def port_lib():
    global _port
    if _port is None:
        _port = C.CDLL(build_port())
    return _port


class Env(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("quat", C.c_double * 4), ("vlin", C.c_double * 3), ("wb", C.c_double * 3),
                ("q", C.c_double * 12), ("qd", C.c_double * 12), ("mass", C.c_double * 13), ("com", C.c_double * 3 * 13),
                ("mu", C.c_double * 2), ("kscale", C.c_double * 2), ("cscale", C.c_double * 2)]


class Terrain(C.Structure):
    _fields_ = [("hf", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("border_pixels", C.c_int),
                ("horizontal_scale", C.c_float), ("vertical_scale", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.t1o_terrain_height.restype = C.c_double
        _lib.t1o_terrain_height.argtypes = [C.c_void_p, C.c_float, C.c_float]
        assert _lib.t1o_sizeof_env() == C.sizeof(Env)
    return _lib


def make_env(model, pos=(0, 0, 0.72), quat=(0, 0, 0, 1), vlin=(0, 0, 0), wb=(0, 0, 0), q=None, qd=None, mu=1.0):
    e = Env()
    e.pos[:] = pos
    e.quat[:] = quat
    e.vlin[:] = vlin
    e.wb[:] = wb
    e.q[:] = q if q is not None else [0.0] * 12
    e.qd[:] = qd if qd is not None else [0.0] * 12
    for b in range(13):
        e.mass[b] = model.mass[b]
        for r in range(3):
            e.com[b][r] = model.ipos[b][r]
    e.mu[:] = [mu, mu]
    e.kscale[:] = [1.0, 1.0]
    e.cscale[:] = [1.0, 1.0]
    return e


def make_terrain(hf=None, border_pixels=50, hscale=0.1, vscale=0.005):
    t = Terrain()
    if hf is None:
        t.hf = None
        t.rows = t.cols = 0
    else:
        hf = np.ascontiguousarray(hf, dtype=np.int16)
        t._keep = hf
        t.hf = hf.ctypes.data
        t.rows, t.cols = hf.shape
    t.border_pixels = border_pixels
    t.horizontal_scale = hscale
    t.vertical_scale = vscale
    return t


def _d(a):
    return (C.c_double * len(a))(*a)


def tick(model, env, tau=None, push_f=(0, 0, 0), push_t=(0, 0, 0), terrain=None, integrate=True):
    """returns (status, qacc[18], foot_fn[2]); env is advanced in place if integrate"""
    terrain = terrain or make_terrain()
    qacc = (C.c_double * 18)()
    fn = (C.c_double * 2)()
    st = lib().t1o_tick(C.byref(model), C.byref(env), _d(tau if tau is not None else [0.0] * 12), _d(push_f), _d(push_t),
                        C.byref(terrain), qacc, fn, 1 if integrate else 0)
    return st, np.array(qacc), np.array(fn)


def tick_f(model, env, tau=None, push_f=(0, 0, 0), push_t=(0, 0, 0), terrain=None, integrate=True):
    """like tick(), also returning the net contact force per body [13, 3] (gym.refresh_net_contact_force_tensor rows)"""
    terrain = terrain or make_terrain()
    qacc = (C.c_double * 18)()
    fn = (C.c_double * 2)()
    bf = (C.c_double * 39)()
    st = lib().t1o_tick_f(C.byref(model), C.byref(env), _d(tau if tau is not None else [0.0] * 12), _d(push_f), _d(push_t),
                          C.byref(terrain), qacc, fn, bf, 1 if integrate else 0)
    return st, np.array(qacc), np.array(fn), np.array(bf).reshape(13, 3)


def mass_matrix(model, env):
    M = np.zeros((18, 18))
    lib().t1o_mass_matrix(C.byref(model), C.byref(env), M.ctypes.data_as(C.c_void_p))
    return M


def energy_momentum(model, env):
    E = C.c_double()
    P = (C.c_double * 3)()
    L = (C.c_double * 3)()
    lib().t1o_energy_momentum(C.byref(model), C.byref(env), C.byref(E), P, L)
    return E.value, np.array(P), np.array(L)


def feet(model, env):
    p = np.zeros((2, 3))
    R = np.zeros((2, 3, 3))
    lib().t1o_feet(C.byref(model), C.byref(env), p.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    return p, R


def terrain_height(terrain, x, y):
    return lib().t1o_terrain_height(C.byref(terrain), C.c_float(x), C.c_float(y))
