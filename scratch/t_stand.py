import numpy as np, sys
sys.path.insert(0,'/root/repo')
from booster_gym_b200 import robot as bm
from oracle import physics as op
md = bm.model_d()
q0 = np.array([-0.2,0,0,0.4,-0.25,0]*2); kp=np.array([200.,200,200,200,float(sys.argv[3]) if len(sys.argv)>3 else 50,float(sys.argv[3]) if len(sys.argv)>3 else 50]*2); kd=np.array([5.,5,5,5,1,1]*2)
lim=np.array([45.,30,30,60,24,15]*2)
z0 = float(sys.argv[1]) if len(sys.argv)>1 else 0.72
e = op.make_env(md, pos=(0,0,z0), q=q0, mu=float(sys.argv[2]) if len(sys.argv)>2 else 1.0)
for i in range(5001):
    tau = np.clip(kp*(q0-np.array(e.q[:]))-kd*np.array(e.qd[:]), -lim, lim)
    st,qa,fn = op.tick(md,e,tau)
    if i%500==0:
        p,R = op.feet(md,e)
        print(i, 'z',round(e.pos[2],4),'x',round(e.pos[0],4),'y',round(e.pos[1],4),'quat',np.round(e.quat[:],3),'fn',np.round(fn,1),'footz',np.round(p[:,2],4), st)
