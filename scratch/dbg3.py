import sys, copy, yaml, torch
sys.path.insert(0, '.')
from booster_gym_b200.learner import Learner
from oracle import learner as L
cfg = yaml.safe_load(open('envs/T1.yaml'))
T, N = 24, 4096
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
res = {}
for dt in (torch.float32, torch.float64):
    sdd = {k: v.to(dt).clone() for k, v in sd.items()}
    bufd = {k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in buf.items()}
    omu, osig, olp = L.old_dist(sdd, bufd["obses"], bufd["actions"])
    adam = L.new_adam(sdd); outs = []; lr = 1e-3
    for ep in range(2):
        o = L.epoch(sdd, adam, bufd, last_obs.to(dt), last_priv.to(dt), omu, osig, olp, lr); lr = o["lr"]; outs.append(o)
    res[dt] = outs
for ep in range(2):
    lrn.epoch_a(dev["rewards"], d8, t8, last_obs.cuda(), last_priv.cuda())
    lrn.epoch_b(dev["actions"])
    o32, o64 = res[torch.float32][ep], res[torch.float64][ep]
    g = lrn.views(lrn.grads)
    print("epoch", ep)
    for name in o64["grads"]:
        ours = g[name].cpu().double().reshape(o64["grads"][name].shape)
        f64 = o64["grads"][name]; f32 = o32["grads"][name].double()
        sc = f64.abs().max().item()
        print("  %-18s scale %.3e  ours_rel %.2e  f32_rel %.2e" % (name, sc, (ours - f64).abs().max().item() / sc, (f32 - f64).abs().max().item() / sc))
    for nm, idx, shape in (("values", 0, (T, N)), ("adv", 1, (T, N)), ("mu", 3, (T, N, 12))):
        ours = lrn.buffer(idx, shape).cpu().double(); key = {"values": "values", "adv": "adv_raw", "mu": "mu"}[nm]
        sc = o64[key].abs().max().item()
        print("  %-18s scale %.3e  ours_rel %.2e  f32_rel %.2e" % (nm, sc, (ours - o64[key]).abs().max().item() / sc, (o32[key].double() - o64[key]).abs().max().item() / sc))
    dmu = lrn.buffer(6, (T, N, 12)).cpu().double()
    sc = o64["dmu"].abs().max().item()
    print("  dmu elementwise: ours_rel %.2e f32_rel %.2e" % ((dmu - o64["dmu"]).abs().max().item() / sc, (o32["dmu"].double() - o64["dmu"]).abs().max().item() / sc))
    cs = dmu.sum(dim=(0, 1)); ref = o64["grads"]["actor.6.bias"]
    print("  colsum64(our dmu) vs f64 bias grad rel %.2e ; our kernel bias vs colsum64(our dmu) rel %.2e" % ((cs - ref).abs().max().item() / ref.abs().max().item(), (g["actor.6.bias"].cpu().double() - cs).abs().max().item() / ref.abs().max().item()))
    cs32 = o32["dmu"].double().sum(dim=(0, 1))
    print("  colsum64(f32 dmu) vs f64 rel %.2e" % ((cs32 - ref).abs().max().item() / ref.abs().max().item()))
    e = (dmu - o64["dmu"]); print("  mean signed err per dim / scale", (e.mean(dim=(0, 1)) * T * N / ref.abs().max().item()).tolist()[:4])
    a_ours = lrn.buffer(1, (T, N)).cpu().double(); print("  adv mean ours %.9e f64 %.9e f32 %.9e ; std ours %.9e f64 %.9e" % (a_ours.mean().item(), o64["adv_raw"].mean().item(), o32["adv_raw"].double().mean().item(), a_ours.std().item(), o64["adv_raw"].std().item()))
    print("  SC adv mean/std", lrn.scalars[7].item(), lrn.scalars[8].item())
    lrn.apply()
