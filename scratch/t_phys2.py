import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo')
from booster_gym_b200 import robot as bm
from oracle import physics as op
md = bm.model_d()
# free fall
e = op.make_env(md, pos=(0,0,2.0)); st, qa, _ = op.tick(md, e, integrate=False); print('freefall qacc', qa[:6], np.abs(qa[6:]).max())
# momentum / energy drift in flight, no gravity torque... energy
md2 = bm.model_d(enable_contact=False, enable_limits=False)
rng=np.random.default_rng(1)
e = op.make_env(md2, pos=(0,0,5.0), vlin=rng.normal(size=3), wb=rng.normal(size=3), q=rng.uniform(-0.3,0.3,12), qd=rng.normal(size=12))
E0,P0,L0 = op.energy_momentum(md2,e)
for i in range(500): op.tick(md2,e)
E1,P1,L1 = op.energy_momentum(md2,e)
print('E', E0,E1,'P',P0,P1,'L about origin',L0,L1)
# standing with PD
q0 = np.array([-0.2,0,0,0.4,-0.25,0]*2); kp=np.array([200,200,200,200,50,50]*2.); kd=np.array([5,5,5,5,1,1]*2.)
lim=np.array([45,30,30,60,24,15]*2.)
e = op.make_env(md, pos=(0,0,0.72), q=q0)
for i in range(2500):
    tau = np.clip(kp*(q0-np.array(e.q[:]))-kd*np.array(e.qd[:]), -lim, lim)
    st,qa,fn = op.tick(md,e,tau)
    if i%250==0: print(i, 'z',round(e.pos[2],4),'x',round(e.pos[0],4),'quat',np.round(e.quat[:],4),'fn',np.round(fn,1), st)
