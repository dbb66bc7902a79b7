import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo')
from booster_gym_b200 import robot as bm, _abi
from oracle import physics as op
hc = C.CDLL('/root/repo/tests/hostcheck/libhostcheck.so')
md = bm.model_d(); mf = bm.model_f()
rng = np.random.default_rng(0)
def rq():
    q = rng.normal(size=4); return q/np.linalg.norm(q)
def hc_tick_d(env, tau, pf, pt, integrate=False):
    e2 = op.Env.from_buffer_copy(env)
    qacc = (C.c_double*18)(); fn=(C.c_double*2)()
    hc.hc_tick_d(C.byref(md), C.byref(e2), op._d(tau), op._d(pf), op._d(pt), None, 0,0,50, C.c_float(0.1), C.c_double(0.005), qacc, fn, int(integrate))
    return e2, np.array(qacc), np.array(fn)
worst=0
for it in range(200):
    z = 2.0 if it%2==0 else 0.70
    e = op.make_env(md, pos=(rng.normal(), rng.normal(), z), quat=rq() if it%2==0 else (0,0,0,1), vlin=rng.normal(size=3), wb=rng.normal(size=3)*2,
                    q=rng.uniform(-0.5,0.5,size=12), qd=rng.normal(size=12)*3)
    for b in range(13):
        e.mass[b] *= rng.uniform(0.8,1.2)
        for r in range(3): e.com[b][r] += rng.uniform(-0.01,0.01)
    tau = rng.normal(size=12)*10; pf = rng.normal(size=3)*10; pt = rng.normal(size=3)*2
    st, qa_o, fn_o = op.tick(md, op.Env.from_buffer_copy(e), tau, pf, pt, integrate=False)
    _, qa_h, fn_h = hc_tick_d(e, tau, pf, pt)
    err = np.max(np.abs(qa_o-qa_h))/max(1,np.max(np.abs(qa_o)))
    worst=max(worst,err)
    if it<4: print(it, st, err, fn_o, fn_h, qa_o[:6])
print('worst rel err', worst)
