"""apply_randomization / discount_values / surrogate_loss with the reference's signatures (utils/utils.py:5-52).

`discount_values` runs the CUDA GAE kernel (b200_gae); inside the training loop the same kernel is launched by
b200_ppo_epoch_a together with the time-out bootstrap.  `apply_randomization` and `surrogate_loss` are host-level
helpers kept for API compatibility - in the hot path their arithmetic lives inside the env / loss kernels.
"""
import numpy as np
import torch


def _draw(like, gaussian):
    """one raw sample shaped like `like`: torch RNG for tensors, NumPy's global RNG for Python scalars"""
    if isinstance(like, torch.Tensor):
        return torch.randn_like(like) if gaussian else torch.rand_like(like)
    return np.random.randn() if gaussian else np.random.rand()


def apply_randomization(tensor, params, return_noise=False):
    """params None -> identity; {"distribution": gaussian|uniform, "operation": additive|scaling, "range": [a, b]}.
    gaussian draws a + b * randn (b acts as a standard deviation), uniform a + (b - a) * rand.  With return_noise the RAW
    sample is returned too (that is what ends up in base_mass_scaled, SURVEY 8a note 5)."""
    if params is None:
        return tensor
    dist, op = params["distribution"], params["operation"]
    if dist not in ("gaussian", "uniform"):
        raise ValueError(f"Invalid randomization distribution: {dist}")
    a, b = params["range"]
    noise = _draw(tensor, dist == "gaussian")
    value = a + b * noise if dist == "gaussian" else a + (b - a) * noise
    if op == "additive":
        result = tensor + value
    elif op == "scaling":
        result = tensor * value
    else:
        raise ValueError(f"Invalid randomization operation: {op}")
    return (result, noise) if return_noise else result


def discount_values(rewards, dones, values, last_values, gamma, lam):
    """GAE advantages [T, N] (utils/utils.py:33-44) on the device"""
    from ..learner import gae

    r = rewards.contiguous().float()
    d = dones.contiguous()
    d = d.view(torch.uint8) if d.dtype == torch.bool else d.to(torch.uint8)
    no_timeouts = torch.zeros_like(d)
    adv, _ = gae(r.clone(), d, no_timeouts, values.contiguous().float(), last_values.contiguous().float(), gamma, lam)
    return adv


def surrogate_loss(old_actions_log_prob, actions_log_prob, advantages, e_clip=0.2):
    ratio = torch.exp(actions_log_prob - old_actions_log_prob)
    unclipped = -advantages * ratio
    clipped = -advantages * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)
    return torch.max(unclipped, clipped).mean()
