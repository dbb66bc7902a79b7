"""CPU oracle of the PPO learner path.  TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py baselines).

A torch-CPU restatement of what the reference computes between rollout end and optimizer step:
  * ActorCritic forward                      utils/model.py:9-36
  * old distribution / log-prob              utils/runner.py:123-125
  * time-out bootstrap, GAE, returns, norm   utils/runner.py:135-145, utils/utils.py:33-44
  * the five loss terms and the total loss   utils/runner.py:146-161, utils/utils.py:47-52
  * backward (autograd), clip_grad_norm_, Adam, KL, learning-rate rule   utils/runner.py:162-180
It is written independently (functional style, explicit formulas instead of torch.distributions) and is PINNED against
the live reference classes by tools/make_golden.py -> tests/golden/learner_*.npz (see tests/test_oracle_pinning.py).
Works in float32 (the reference's dtype) or float64 (ground truth used to bound rounding error).
"""
import math

import torch

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def mlp(x, layers):
    """layers = [(W, b), ...]; ELU between layers, none after the last (utils/model.py:9-26)"""
    for i, (w, b) in enumerate(layers):
        x = x @ w.t() + b
        if i + 1 < len(layers):
            x = elu(x)
    return x


def split_params(sd):
    """state_dict (reference names) -> (critic layers, actor layers, logstd)"""
    critic = [(sd[f"critic.{k}.weight"], sd[f"critic.{k}.bias"]) for k in (0, 2, 4, 6)]
    actor = [(sd[f"actor.{k}.weight"], sd[f"actor.{k}.bias"]) for k in (0, 2, 4, 6)]
    return critic, actor, sd["logstd"]


def init_params(seed=0, dtype=torch.float32, num_act=12, num_obs=47, num_priv=14):
    """torch.nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for W and b),
    logstd = -2 (utils/model.py:27). Deterministic in `seed`; returns a dict with the reference's state_dict names."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(prefix, n_out, n_in):
        bound = 1.0 / math.sqrt(n_in)
        sd[prefix + ".weight"] = ((torch.rand(n_out, n_in, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        sd[prefix + ".bias"] = ((torch.rand(n_out, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    dims_c = [num_obs + num_priv, 256, 256, 128, 1]
    dims_a = [num_obs, 256, 128, 128, num_act]
    for i, k in enumerate((0, 2, 4, 6)):
        lin(f"critic.{k}", dims_c[i + 1], dims_c[i])
    for i, k in enumerate((0, 2, 4, 6)):
        lin(f"actor.{k}", dims_a[i + 1], dims_a[i])
    sd["logstd"] = torch.full((1, num_act), -2.0, dtype=dtype)
    return sd


def actor_mean(sd, obs):
    return mlp(obs, split_params(sd)[1])


def critic_value(sd, obs, priv):
    return mlp(torch.cat((obs, priv), dim=-1), split_params(sd)[0]).squeeze(-1)


def normal_log_prob(x, mu, sigma):
    """torch.distributions.Normal.log_prob: -(x-mu)^2 / (2 sigma^2) - log sigma - log sqrt(2 pi)"""
    return -((x - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - math.log(math.sqrt(2 * math.pi))


def gae(rewards, dones, values, last_values, gamma, lam):
    """utils/utils.py:33-44"""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    for t in range(T - 1, -1, -1):
        nnt = 1.0 - dones[t].to(rewards.dtype)
        nxt = last_values if t == T - 1 else values[t + 1]
        delta = rewards[t] + gamma * nnt * nxt - values[t]
        last = delta + gamma * lam * nnt * last
        adv[t] = last
    return adv


def surrogate(old_logp, logp, adv, e_clip=0.2):
    """utils/utils.py:47-52"""
    ratio = torch.exp(logp - old_logp)
    return torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)).mean()


def old_dist(sd, obses, actions):
    """utils/runner.py:123-125 -> (old_mu [T,N,12], old_sigma [1,12], old_logp [T,N])"""
    with torch.no_grad():
        mu = actor_mean(sd, obses)
        sigma = torch.exp(sd["logstd"])
        logp = normal_log_prob(actions, mu, sigma.expand_as(mu)).sum(dim=-1)
    return mu, sigma, logp


def epoch(sd, adam, buf, last_obs, last_priv, old_mu, old_sigma, old_logp, lr, gamma=0.995, lam=0.95, bound_coef=1.0,
          entropy_coef=-0.01, desired_kl=0.01, max_grad_norm=1.0, betas=(0.9, 0.999), eps=1e-8):
    """One full-batch epoch (utils/runner.py:132-180). `sd` (leaf tensors), `adam` = {"step": int, "m": {..}, "v": {..}}
    and buf["rewards"] are updated in place; returns a dict of every intermediate the CUDA path is compared on."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    obses, privs, actions = buf["obses"], buf["privileged_obses"], buf["actions"]
    values = critic_value(params, obses, privs)
    last_values = critic_value(params, last_obs, last_priv)
    with torch.no_grad():
        buf["rewards"][buf["time_outs"]] = values[buf["time_outs"]]
        adv_raw = gae(buf["rewards"], buf["dones"] | buf["time_outs"], values, last_values, gamma, lam)
        returns = values + adv_raw
        adv = (adv_raw - adv_raw.mean()) / (adv_raw.std() + 1e-8)
    value_loss = ((values - returns) ** 2).mean()
    mu = actor_mean(params, obses)
    mu.retain_grad()
    values.retain_grad()
    sigma = torch.exp(params["logstd"]).expand_as(mu)
    logp = normal_log_prob(actions, mu, sigma).sum(dim=-1)
    actor_loss = surrogate(old_logp, logp, adv)
    bound_loss = torch.clip(mu - 1.0, min=0.0).square().mean() + torch.clip(mu + 1.0, max=0.0).square().mean()
    entropy = (0.5 + HALF_LOG_2PI + torch.log(sigma)).sum(dim=-1)
    loss = value_loss + actor_loss + bound_coef * bound_loss + entropy_coef * entropy.mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in params.items()}
    # clip_grad_norm_(parameters, 1.0)
    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(mu.dtype)
    coef = torch.clamp(max_grad_norm / (total_norm + 1e-6), max=1.0)
    clipped = {k: g * coef for k, g in grads.items()}
    # torch.optim.Adam (defaults)
    adam["step"] += 1
    t = adam["step"]
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    with torch.no_grad():
        for k in sd:
            g = clipped[k]
            m, v = adam["m"][k], adam["v"][k]
            m.lerp_(g, 1 - betas[0])
            v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
            denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
            sd[k].addcdiv_(m, denom, value=-(lr / bc1))
        sig = sigma.detach()
        kl = (torch.log(sig / old_sigma) + 0.5 * (old_sigma ** 2 + (mu.detach() - old_mu) ** 2) / sig ** 2 - 0.5).sum(dim=-1)
        kl_mean = kl.mean()
    new_lr = lr
    if kl_mean > desired_kl * 2:
        new_lr = max(1e-5, lr / 1.5)
    elif kl_mean < desired_kl / 2:
        new_lr = min(1e-2, lr * 1.5)
    return dict(values=values.detach(), last_values=last_values.detach(), adv_raw=adv_raw, returns=returns, adv=adv,
                mu=mu.detach(), logp=logp.detach(), value_loss=value_loss.item(), actor_loss=actor_loss.item(),
                bound_loss=bound_loss.item(), entropy=entropy.mean().item(), kl=kl_mean.item(), grads=grads,
                grad_norm=float(total_norm), lr=new_lr, dmu=mu.grad.detach().clone(), dvalues=values.grad.detach().clone())


def new_adam(sd):
    return {"step": 0, "m": {k: torch.zeros_like(v) for k, v in sd.items()}, "v": {k: torch.zeros_like(v) for k, v in sd.items()}}


def synthetic_rollout(T, N, seed=0, dtype=torch.float32, done_rate=0.002, timeout_rate=0.003):
    """the synthetic buffer used by parity tests and the CPU baseline (BASELINE.md section 3 distribution)"""
    g = torch.Generator().manual_seed(seed)
    buf = {
        "obses": torch.randn(T, N, 47, generator=g).to(dtype),
        "privileged_obses": torch.randn(T, N, 14, generator=g).to(dtype),
        "actions": torch.randn(T, N, 12, generator=g).clamp(-1.5, 1.5).to(dtype) * 0.5,
        "rewards": (torch.rand(T, N, generator=g) * 0.05).to(dtype),
        "dones": torch.rand(T, N, generator=g) < done_rate,
        "time_outs": torch.rand(T, N, generator=g) < timeout_rate,
    }
    last_obs = torch.randn(N, 47, generator=g).to(dtype)
    last_priv = torch.randn(N, 14, generator=g).to(dtype)
    return buf, last_obs, last_priv
