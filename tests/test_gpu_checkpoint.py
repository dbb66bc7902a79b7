"""checkpoint / export compatibility.  The `.pth` dict the Runner writes (`model` / `optimizer` / `curriculum`,
utils/runner.py:206-214, utils/recorder.py:70-73) must load into the REFERENCE's own objects - a plain
`torch.nn` ActorCritic with the reference layout and `torch.optim.Adam(model.parameters())` - and back, and the TorchScript
actor that export_model.py:19-29 produces from it must reproduce the kernel's action means (deploy/utils/policy.py loads it)."""
import copy
import io

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(t1_cfg, N=64, T=4):
    from booster_gym_b200.learner import Learner
    from booster_gym_b200.utils.model import ActorCritic
    from booster_gym_b200.utils.runner import FlatAdam
    from oracle import learner as L

    cfg = copy.deepcopy(t1_cfg)
    cfg["runner"]["horizon_length"] = T
    lrn = Learner(cfg, N, "cuda:0", learning_rate=3e-4, seed=7)
    model = ActorCritic(12, 47, 14)
    torch.manual_seed(3)
    for p in model.parameters():          # non-trivial parameters (default nn.Linear init + perturbed logstd)
        if p.dim() == 2 and p.shape[0] == 1:
            p.data += 0.1 * torch.randn_like(p)
    model.bind(lrn)
    opt = FlatAdam(model, lrn)
    buf, last_obs, last_priv = L.synthetic_rollout(T, N, seed=11)
    dev = {k: v.cuda() for k, v in buf.items()}
    lrn.old_dist(dev["obses"], dev["privileged_obses"], dev["actions"])
    for _ in range(2):                     # two Adam steps so that exp_avg / exp_avg_sq / step are populated
        lrn.epoch_a(dev["rewards"], dev["dones"].to(torch.uint8), dev["time_outs"].to(torch.uint8), last_obs.cuda(), last_priv.cuda())
        lrn.epoch_b(dev["actions"])
        lrn.apply()
    torch.cuda.synchronize()
    return cfg, lrn, model, opt, buf


def test_checkpoint_loads_into_reference_shaped_objects_and_back(t1_cfg, tmp_path):
    from booster_gym_b200 import _abi
    from booster_gym_b200.learner import Learner
    from booster_gym_b200.utils.model import ActorCritic
    from booster_gym_b200.utils.runner import FlatAdam

    cfg, lrn, model, opt, _ = _setup(t1_cfg)
    curriculum = torch.rand(21, 21)
    path = tmp_path / "model_10.pth"
    torch.save({"model": model.state_dict(), "optimizer": opt.state_dict(), "curriculum": curriculum}, path)
    ckpt = torch.load(path, map_location="cpu", weights_only=True)   # the reference loads with weights_only=True (utils/runner.py:88)
    assert set(ckpt) == {"model", "optimizer", "curriculum"}

    # (1) reference-shaped consumers: nn.Module with the reference layout (strict) and a real torch.optim.Adam
    ref = ActorCritic(12, 47, 14)                                    # unbound: a plain CPU torch module, same definition as utils/model.py:5-27
    missing = ref.load_state_dict(ckpt["model"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1.0)
    ref_opt.load_state_dict(ckpt["optimizer"])
    assert ref_opt.param_groups[0]["lr"] == pytest.approx(float(lrn.scalars[_abi.SC["LR"]].item()))
    m = lrn.views(lrn.adam_m)
    for (name, p) in ref.named_parameters():
        st = ref_opt.state[p]
        assert float(st["step"]) == 2.0
        assert torch.equal(st["exp_avg"].reshape(-1), m[name].cpu().reshape(-1)), name
        assert torch.equal(p.detach().reshape(-1), lrn.views()[name].cpu().reshape(-1)), name
    ref_opt.step  # (a torch Adam continues from this state: layout accepted)

    # (2) back into a fresh learner: bit-exact parameters, Adam moments, step counter and learning rate
    lrn2 = Learner(cfg, 64, "cuda:0", learning_rate=1.0, seed=0)
    model2 = ActorCritic(12, 47, 14, learner=lrn2)
    opt2 = FlatAdam(model2, lrn2)
    model2.load_state_dict(ckpt["model"], strict=False)             # utils/runner.py:89
    opt2.load_state_dict(ckpt["optimizer"])
    n = _abi.NPARAMS
    for a, b in ((lrn.params, lrn2.params), (lrn.adam_m, lrn2.adam_m), (lrn.adam_v, lrn2.adam_v)):
        va, vb = lrn.views(a), lrn2.views(b)
        for name in va:
            assert torch.equal(va[name], vb[name]), name
    assert lrn2.scalars[_abi.SC["ADAM_STEP"]].item() == 2.0
    assert lrn2.scalars[_abi.SC["LR"]].item() == lrn.scalars[_abi.SC["LR"]].item()
    assert n == 177945


def test_torchscript_actor_export_matches_the_kernel(t1_cfg):
    """export_model.py:19-29: torch.jit.script(model.actor) saved next to the checkpoint; deploy/utils/policy.py:44-63 runs it"""
    from booster_gym_b200.utils.model import ActorCritic

    cfg, lrn, model, opt, buf = _setup(t1_cfg)
    ref = ActorCritic(12, 47, 14)
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref.eval()
    f = io.BytesIO()
    torch.jit.save(torch.jit.script(ref.actor), f)
    f.seek(0)
    scripted = torch.jit.load(f, map_location="cpu")
    obs = buf["obses"][0]                                            # [64, 47]
    want = scripted(obs)
    dist = model.act(obs.cuda())                                     # b200_policy_act, deterministic mean
    got = dist.loc.cpu()
    assert got.shape == want.shape == (64, 12)
    assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    assert torch.allclose(dist.scale.cpu(), torch.exp(ref.logstd).expand_as(want))
