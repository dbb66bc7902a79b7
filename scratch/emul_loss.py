import sys, torch, numpy as np
sys.path.insert(0, '.')
from oracle import learner as L
T, N = 24, 1024
sd = L.init_params(0); sd["actor.6.weight"] *= 8.0; sd["logstd"] += torch.linspace(-0.3, 0.3, 12).view(1, 12)
buf, lo, lp_ = L.synthetic_rollout(T, N, seed=3, done_rate=0.02, timeout_rate=0.03)
mu0 = L.actor_mean(sd, buf["obses"])
buf["actions"] = mu0 + torch.exp(sd["logstd"]) * torch.randn(T, N, 12, generator=torch.Generator().manual_seed(5))
res = {}
for dt in (torch.float32, torch.float64):
    sdd = {k: v.to(dt).clone() for k, v in sd.items()}
    bufd = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in buf.items()}
    omu, osig, olp = L.old_dist(sdd, bufd["obses"], bufd["actions"])
    res[dt] = L.epoch(sdd, L.new_adam(sdd), bufd, lo.to(dt), lp_.to(dt), omu, osig, olp, 1e-3)
o32, o64 = res[torch.float32], res[torch.float64]
M = T * N
# emulate k_loss in fp32 with f32 oracle inputs
f = np.float32
adv = o32["adv_raw"].numpy().astype(np.float64)
mean = adv.mean(); std = adv.std(ddof=1)
A = ((o32["adv_raw"].numpy() - f(mean)) / (f(std) + f(1e-8))).astype(np.float32)
mu = o32["mu"].numpy(); a = buf["actions"].numpy()
sg = np.exp(sd["logstd"].numpy()).astype(np.float32)[0]
var = sg * sg
invM = f(1.0) / f(M)
dlp = (-A * f(1.0) * invM).astype(np.float32)
d = a - mu
up = np.maximum(mu - f(1), f(0)); dn = np.minimum(mu + f(1), f(0))
bscale = f(1.0) * f(2.0) / (f(M) * f(12.0))
dmu = (dlp[..., None] * d / var + bscale * (up + dn)).astype(np.float32)
ref = o64["grads"]["actor.6.bias"].numpy()
sc = np.abs(ref).max()
print("emulated colsum rel err", np.abs(dmu.astype(np.float64).sum((0, 1)) - ref).max() / sc)
print("f32 autograd dmu colsum rel err", np.abs(o32["dmu"].double().numpy().sum((0, 1)) - ref).max() / sc)
# variants
A64 = (o64["adv_raw"].numpy() - o64["adv_raw"].numpy().mean()) / (o64["adv_raw"].numpy().std(ddof=1) + 1e-8)
for name, Ause in (("A from f64", A64.astype(np.float32)), ("A f32-torch", o32["adv"].numpy())):
    dlp = (-Ause * invM).astype(np.float32)
    dm = (dlp[..., None] * d / var + bscale * (up + dn)).astype(np.float32)
    print(name, np.abs(dm.astype(np.float64).sum((0, 1)) - ref).max() / sc)
print("A mean f32-torch", o32["adv"].double().mean().item(), "A emul mean", A.astype(np.float64).mean(), "A64 mean", A64.mean())
