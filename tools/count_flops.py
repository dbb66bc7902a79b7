"""Exact floating-point operation count of ONE physics tick of booster_gym_b200/csrc/t1_dynamics.cuh (the recursion k_physics
runs), by instantiating the tick template with a counting scalar (tools/count_flops.cpp).  Replaces SURVEY 8(d)'s survey-time
estimate of 1.28e4 FLOP per env-substep; the numbers are frozen in BASELINE.md section 5.

    python tools/count_flops.py
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from booster_gym_b200 import robot
    from oracle import physics as op

    so = os.path.join(ROOT, "build", "libcountflops.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "count_flops.cpp")])
    lib = C.CDLL(so)
    md = robot.model_d()
    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2)
    rng = np.random.default_rng(0)
    out = {}
    cases = {
        "contact_free (base z = 2.0 m, random joint state, PD torques)": dict(pos=(0.0, 0.0, 2.0), q=q0 + rng.normal(0, 0.2, 12), qd=rng.normal(0, 1.0, 12)),
        "full_foot_contact (standing, both soles loaded: 8 corner contacts)": dict(pos=(0.0, 0.0, 0.66), q=q0, qd=np.zeros(12)),
    }
    for name, kw in cases.items():
        env = op.make_env(md, **kw)
        tau = (C.c_double * 12)(*(rng.normal(0, 5.0, 12)))
        counts = (C.c_longlong * 5)()
        fn = (C.c_double * 2)()
        lib.cf_tick(C.byref(md), C.byref(env), tau, counts, fn)
        a, m, d, s, c = [int(x) for x in counts]
        out[name] = {"add_sub": a, "mul": m, "div": d, "sqrt_rsqrt_sin_cos": s, "compare_abs_max": c,
                     "flop": a + m + d + s, "flop_with_compares": a + m + d + s + c, "foot_normal_force_N": [fn[0], fn[1]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
