"""ActorCritic (mirror of utils/model.py:5-36) whose parameters live in the learner's flat CUDA buffer.

Parameter names, shapes and `state_dict()` layout equal the reference module (`critic.{0,2,4,6}.*`, `actor.{0,2,4,6}.*`,
`logstd[1,12]`), so checkpoints interchange with play_mujoco.py / export_model.py.  `act()` / `est_value()` run the
tensor-core MLP kernels (b200_policy_act / b200_critic_value) and are inference-only: training gradients are produced
by b200_ppo_epoch, not by autograd.  The torch modules are containers for the parameter views (and give a CPU-callable
actor for export), they are not what computes on the hot path.
"""
import torch

from ..learner import Learner


class ActorCritic(torch.nn.Module):

    def __init__(self, num_act, num_obs, num_privileged_obs, learner=None):
        super().__init__()
        if (num_act, num_obs, num_privileged_obs) != (12, 47, 14):
            raise ValueError("the MLP kernels are built for ActorCritic(12, 47, 14)")
        self.critic = torch.nn.Sequential(
            torch.nn.Linear(num_obs + num_privileged_obs, 256), torch.nn.ELU(),
            torch.nn.Linear(256, 256), torch.nn.ELU(),
            torch.nn.Linear(256, 128), torch.nn.ELU(),
            torch.nn.Linear(128, 1),
        )
        self.actor = torch.nn.Sequential(
            torch.nn.Linear(num_obs, 256), torch.nn.ELU(),
            torch.nn.Linear(256, 128), torch.nn.ELU(),
            torch.nn.Linear(128, 128), torch.nn.ELU(),
            torch.nn.Linear(128, num_act),
        )
        self.logstd = torch.nn.parameter.Parameter(torch.full((1, num_act), fill_value=-2.0), requires_grad=True)
        self._learner = None
        if learner is not None:
            self.bind(learner)

    def bind(self, learner: Learner):
        """move the parameters into the learner's flat buffer (keeps current values) and alias .grad to its gradient buffer"""
        views, gviews = learner.views(), learner.views(learner.grads)
        for name, p in list(self.named_parameters()):
            views[name].copy_(p.detach().to(views[name].device).reshape(views[name].shape))
            mod = self
            parts = name.split(".")
            for part in parts[:-1]:
                mod = getattr(mod, part)
            newp = torch.nn.Parameter(views[name], requires_grad=True)
            newp.grad = gviews[name]
            setattr(mod, parts[-1], newp)
        self._learner = learner
        return self

    def to(self, *args, **kwargs):
        if self._learner is not None:
            return self  # parameters already live on the learner's device; moving them would break the aliasing
        return super().to(*args, **kwargs)

    def _need(self):
        if self._learner is None:
            raise RuntimeError("ActorCritic is not bound to a Learner (no CPU path exists in this package)")
        return self._learner

    @torch.no_grad()
    def act(self, obs):
        lrn = self._need()
        flat = obs.reshape(-1, obs.shape[-1]).contiguous().float()
        mu = torch.empty(flat.shape[0], 12, dtype=torch.float32, device=flat.device)
        chunk = lrn.num_envs
        for s in range(0, flat.shape[0], chunk):
            part = flat[s:s + chunk]
            lrn.act(part, mu[s:s + chunk], deterministic=True, step=0)
        action_mean = mu.view(*obs.shape[:-1], 12)
        action_std = torch.exp(self.logstd).expand_as(action_mean)
        return torch.distributions.Normal(action_mean, action_std)

    @torch.no_grad()
    def est_value(self, obs, privileged_obs):
        lrn = self._need()
        o = obs.reshape(-1, obs.shape[-1]).contiguous().float()
        p = privileged_obs.reshape(-1, privileged_obs.shape[-1]).contiguous().float()
        out = torch.empty(o.shape[0], dtype=torch.float32, device=o.device)
        chunk = lrn.num_envs
        for s in range(0, o.shape[0], chunk):
            lrn.value(o[s:s + chunk], p[s:s + chunk], out[s:s + chunk])
        return out.view(obs.shape[:-1])
