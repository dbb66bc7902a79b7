import sys, copy, yaml, torch
sys.path.insert(0, '.')
from booster_gym_b200.learner import Learner
from oracle import learner as L
cfg = yaml.safe_load(open('envs/T1.yaml'))
T, N = 3, 200
cfg["runner"]["horizon_length"] = T
lrn = Learner(cfg, N, "cuda:0", learning_rate=1e-3)
sd = L.init_params(0); sd["actor.6.weight"] *= 8.0; sd["logstd"] += torch.linspace(-0.3, 0.3, 12).view(1, 12)
lrn.load_state_dict(sd)
buf, lo, lp = L.synthetic_rollout(T, N, seed=3)
mu0 = L.actor_mean(sd, buf["obses"])
buf["actions"] = mu0 + torch.exp(sd["logstd"]) * torch.randn(T, N, 12, generator=torch.Generator().manual_seed(5))
dev = {k: v.cuda() for k, v in buf.items()}
lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
sd64 = {k: v.double() for k, v in sd.items()}
omu, osig, olp = L.old_dist(sd64, buf["obses"].double(), buf["actions"].double())
omu32, osig32, olp32 = L.old_dist(sd, buf["obses"], buf["actions"])
e_mu = (lrn.old_mu.cpu().view(T, N, 12).double() - omu).abs()
e_lp = (lrn.old_logp.cpu().view(T, N).double() - olp).abs()
print('mu err max', e_mu.max().item(), 'argmax', e_mu.flatten().argmax().item(), 'scale', omu.abs().max().item(), 'f32 err', (omu32.double() - omu).abs().max().item())
print('lp err max', e_lp.max().item(), 'argmax', e_lp.flatten().argmax().item(), 'mean', e_lp.mean().item(), 'f32 err', (olp32.double() - olp).abs().max().item(), (olp32.double() - olp).abs().mean().item())
print('sorted top lp errs', e_lp.flatten().sort(descending=True).values[:8])
# direct: logp computed on CPU in fp32 from the GPU mu
mu_g = lrn.old_mu.cpu().view(T, N, 12)
lp_from_gmu = L.normal_log_prob(buf["actions"], mu_g, torch.exp(sd["logstd"]).expand_as(mu_g)).sum(-1)
print('gpu logp vs cpu-logp-of-gpu-mu', (lrn.old_logp.cpu().view(T, N) - lp_from_gmu).abs().max().item())
