import sys, copy, yaml, torch
sys.path.insert(0, '.')
from booster_gym_b200.learner import Learner
from booster_gym_b200 import _abi
from oracle import learner as L
cfg = yaml.safe_load(open('envs/T1.yaml'))
for T, N in ((3, 200), (24, 4096)):
    cfg["runner"]["horizon_length"] = T
    lrn = Learner(cfg, N, "cuda:0", learning_rate=1e-3)
    sd = L.init_params(0); sd["actor.6.weight"] *= 8.0; sd["logstd"] += torch.linspace(-0.3, 0.3, 12).view(1, 12)
    lrn.load_state_dict(sd)
    buf, last_obs, last_priv = L.synthetic_rollout(T, N, seed=3, done_rate=0.02, timeout_rate=0.03)
    mu0 = L.actor_mean(sd, buf["obses"])
    buf["actions"] = mu0 + torch.exp(sd["logstd"]) * torch.randn(T, N, 12, generator=torch.Generator().manual_seed(5))
    dev = {k: v.cuda() for k, v in buf.items()}
    d8, t8 = dev["dones"].to(torch.uint8), dev["time_outs"].to(torch.uint8)
    lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
    sdd = {k: v.double().clone() for k, v in sd.items()}
    bufd = {k: (v.double().clone() if v.is_floating_point() else v.clone()) for k, v in buf.items()}
    omu, osig, olp = L.old_dist(sdd, bufd["obses"], bufd["actions"])
    adam = L.new_adam(sdd); lr = 1e-3
    for ep in range(2):
        o = L.epoch(sdd, adam, bufd, last_obs.double(), last_priv.double(), omu, osig, olp, lr); lr = o["lr"]
        lrn.epoch_a(dev["rewards"], d8, t8, last_obs.cuda(), last_priv.cuda()); lrn.epoch_b(dev["actions"])
        g = lrn.views(lrn.grads)
        tot = 0.0
        for name, ref in o["grads"].items():
            ours = g[name].cpu().double().reshape(ref.shape)
            n_o, n_r = ours.norm().item(), ref.norm().item()
            print(f"T{T} ep{ep} {name:18s} |ours| {n_o:.6e} |ref| {n_r:.6e} rel diff of norms {(n_o-n_r)/n_r:+.2e}  max err/max {((ours-ref).abs().max()/ref.abs().max()).item():.2e}")
        lrn.apply()
        print("   grad norm ours", lrn.scalars[_abi.SC["GRAD_NORM"]].item(), "ref", o["grad_norm"])
