"""Host-side logic that needs no GPU: YAML -> C config, terrain generation, env-origin layout and sharding, buffers,
checkpoint-compatible Adam facade layout, apply_randomization, multi-process (gloo, world_size 2) gradient averaging."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_builder_matches_yaml(t1_cfg):
    from booster_gym_b200 import _abi, config

    c = config.t1_config(t1_cfg)
    assert abs(c.env_dt - 0.02) < 1e-9 and c.decimation == 10
    assert c.n_rew == 23  # dof_vel_limits, torque_limits, feet_vel_z have scale -0. and are dropped (envs/t1.py:279-285)
    names = [_abi.REW_NAMES[c.rew_id[k]] for k in range(c.n_rew)]
    assert names == [k for k, v in t1_cfg["rewards"]["scales"].items() if v != 0]
    assert np.isclose(c.rew_scale[0], 0.25 * 0.02) and np.isclose(c.rew_scale[4], -20.0 * 0.02)
    assert (c.kick_interval, c.push_interval, c.push_duration) == (100, 250, 50)
    assert (c.resample_lo, c.resample_hi, c.max_episode_length) == (400, 600, 1500)
    assert list(c.torque_limits) == [45.0, 30.0, 30.0, 60.0, 24.0, 15.0] * 2   # URDF efforts (not the MJCF ctrlrange)
    assert np.allclose(list(c.default_dof_pos), [-0.2, 0, 0, 0.4, -0.25, 0] * 2)
    assert np.allclose(list(c.kp_nominal), [200, 200, 200, 200, 50, 50] * 2) and np.allclose(list(c.kd_nominal), [5, 5, 5, 5, 1, 1] * 2)
    assert bin(c.penalized_body_mask).count("1") == 11 and not (c.penalized_body_mask & (1 << 6 | 1 << 12))
    assert c.terrain_type == 1 and c.border_pixels == 50 and c.vertical_scale == 0.005
    assert c.init_dof_pos.enabled and c.init_dof_pos.dist == 0 and c.init_base_pos_xy.dist == 1 and np.isclose(c.init_base_pos_xy.b, 2.0)
    p = config.ppo_config(t1_cfg, 4096, world_size=8, env_base=4096)
    assert (p.gamma, p.lam, p.horizon, p.num_envs, p.world_size, p.env_base) == (0.995, 0.95, 24, 4096, 8, 4096)
    assert np.isclose(p.entropy_coef, -0.01) and np.isclose(p.e_clip, 0.2)


def test_effort_limit_source_and_build_flag_hash(t1_cfg):
    """asset.effort_limits: "mjcf" selects the MJCF actuator ctrlrange (play_mujoco.py:753-755, SURVEY 8a note 10) instead of the URDF efforts;
    object files of the in-tree build are named after their flags, so a -D measurement build can never be re-linked under the default stamp"""
    import hashlib

    from booster_gym_b200 import _build, config

    cfg = copy.deepcopy(t1_cfg)
    cfg.setdefault("asset", {})["effort_limits"] = "mjcf"
    assert list(config.t1_config(cfg).torque_limits) == [45.0, 45.0, 30.0, 65.0, 24.0, 15.0] * 2
    tag = hashlib.sha256(" ".join(_build.ARCH + _build.COMMON + _build.UNITS["env_kernels.cu"]).encode()).hexdigest()[:10]
    assert tag != hashlib.sha256(" ".join(_build.ARCH + _build.COMMON + ["-DX"] + _build.UNITS["env_kernels.cu"]).encode()).hexdigest()[:10]
    if os.path.isdir(_build.OBJ_DIR) and os.path.exists(_build.LIB) and _build.up_to_date():
        assert os.path.exists(os.path.join(_build.OBJ_DIR, "env_kernels." + tag + ".o"))


def test_config_errors_follow_the_reference(t1_cfg):
    from booster_gym_b200 import config

    bad = copy.deepcopy(t1_cfg)
    bad["terrain"]["type"] = "lava"
    with pytest.raises(ValueError, match="Invalid terrain type"):
        config.t1_config(bad)
    bad = copy.deepcopy(t1_cfg)
    bad["noise"]["gravity"]["distribution"] = "cauchy"
    with pytest.raises(ValueError, match="Invalid randomization distribution"):
        config.t1_config(bad)
    bad = copy.deepcopy(t1_cfg)
    bad["randomization"]["base_mass"]["operation"] = "power"
    with pytest.raises(ValueError, match="Invalid randomization operation"):
        config.t1_config(bad)
    bad = copy.deepcopy(t1_cfg)
    bad["control"]["stiffness"] = {"Hip": 1.0}
    with pytest.raises(ValueError, match="PD gain"):
        config.t1_config(bad)
    none = copy.deepcopy(t1_cfg)
    none["randomization"]["kick_lin_vel"] = None
    del none["noise"]["height"]
    c = config.t1_config(none)
    assert c.kick_lin_vel.enabled == 0 and c.noise_height.enabled == 0  # apply_randomization(x, None) is the identity


def test_terrain_generation_and_layout(t1_cfg):
    from booster_gym_b200.utils.terrain import Terrain

    np.random.seed(42)
    t = Terrain(None, None, "cpu", t1_cfg["terrain"])
    hf = t.height_field_raw
    assert hf.shape == (900, 200) and hf.dtype == np.int16 and t.border_pixels == 50
    assert (t.env_width, t.env_length, t.border_size) == (80.0, 10.0, 5.0)
    assert not hf[:50].any() and not hf[-50:].any() and not hf[:, :50].any() and not hf[:, -50:].any()  # flat border
    rough, discrete = hf[50:450, 50:150], hf[450:850, 50:150]
    assert np.abs(rough).max() <= 10 and len(np.unique(rough)) > 5           # +-0.05 m in 5 mm units
    assert set(np.unique(discrete)) <= {-4, -2, 0, 2, 4}                      # {-h, -h//2, h//2, h} with h = 0.02 / 0.005
    np.random.seed(42)
    assert np.array_equal(Terrain(None, None, "cpu", t1_cfg["terrain"]).height_field_raw, hf)  # seeded
    plane = copy.deepcopy(t1_cfg["terrain"]); plane["type"] = "plane"
    tp = Terrain(None, None, "cpu", plane)
    assert tp.height_field_raw is None and torch.equal(tp.terrain_heights(torch.zeros(5, 3)), torch.zeros(5))
    with pytest.raises(ValueError):
        Terrain(None, None, "cpu", dict(plane, type="moon"))


def test_env_origins_are_a_global_layout_shared_by_all_ranks(t1_cfg):
    """rank r owns envs [r*N, (r+1)*N) of the layout a single process would build for world*N envs"""
    from booster_gym_b200 import config

    cfg = copy.deepcopy(t1_cfg)
    cfg["terrain"]["type"] = "plane"
    full = config.env_origins(cfg, 8 * 64)
    assert full.shape == (512, 3)
    cols = int(np.floor(np.sqrt(512)))
    assert np.allclose(full[cols + 1, :2], [1.0, 1.0]) and np.allclose(full[1, :2], [0.0, 1.0])
    cfg["terrain"]["type"] = "trimesh"
    o = config.env_origins(cfg, 4096, lambda xyz: np.full(len(xyz), 0.25, np.float32))
    assert o[:, 0].min() > 0 and o[:, 0].max() < 80 and o[:, 1].min() > 0 and o[:, 1].max() < 10 and np.all(o[:, 2] == 0.25)
    cols = max(1.0, np.floor(np.sqrt(4096 * 10 / 80)))
    assert np.isclose(o[1, 1] - o[0, 1], 10.0 / (cols + 1))


def test_experience_buffer_interface():
    from booster_gym_b200.utils.buffer import ExperienceBuffer

    b = ExperienceBuffer(3, 5, "cpu")
    b.add_buffer("obses", (47,))
    b.add_buffer("rewards", ())
    b.add_buffer("dones", (), dtype=bool)
    assert len(b) == 3 and set(b.keys()) == {"obses", "rewards", "dones"}
    assert b["obses"].shape == (3, 5, 47) and b["dones"].dtype == torch.bool and not b["obses"].any()
    b.update_data("rewards", 1, torch.arange(5.0))
    b.update_data("dones", 2, torch.tensor([True, False, True, False, False]))
    assert torch.equal(b["rewards"][1], torch.arange(5.0)) and b["rewards"][0].sum() == 0
    raw = b.raw("dones")
    assert raw.dtype == torch.uint8 and raw.data_ptr() == b["dones"].data_ptr() and raw[2].tolist() == [1, 0, 1, 0, 0]
    assert b.row("rewards", 1).is_contiguous()


def test_apply_randomization_semantics():
    from booster_gym_b200.utils.utils import apply_randomization, surrogate_loss

    x = torch.ones(1000, 3)
    assert apply_randomization(x, None) is x
    torch.manual_seed(0)
    y, raw = apply_randomization(x, {"distribution": "uniform", "operation": "scaling", "range": [0.8, 1.2]}, return_noise=True)
    assert raw.min() >= 0 and raw.max() < 1 and torch.allclose(y, x * (0.8 + 0.4 * raw))       # RAW sample is returned
    y, raw = apply_randomization(x, {"distribution": "gaussian", "operation": "additive", "range": [0.5, 0.1]}, return_noise=True)
    assert torch.allclose(y, x + 0.5 + 0.1 * raw) and abs(raw.std().item() - 1.0) < 0.1          # "var" acts as a std
    assert isinstance(apply_randomization(2.0, {"distribution": "uniform", "operation": "additive", "range": [0.0, 1.0]}), float)
    with pytest.raises(ValueError):
        apply_randomization(x, {"distribution": "beta", "operation": "additive", "range": [0, 1]})
    with pytest.raises(ValueError):
        apply_randomization(x, {"distribution": "uniform", "operation": "xor", "range": [0, 1]})
    from oracle import learner as L

    a, b, adv = torch.randn(100), torch.randn(100), torch.randn(100)
    assert torch.allclose(surrogate_loss(a, b, adv), L.surrogate(a, b, adv))


def test_actor_critic_module_has_the_reference_state_dict_layout():
    from booster_gym_b200.utils.model import ActorCritic

    m = ActorCritic(12, 47, 14)
    keys = list(m.state_dict().keys())
    assert keys == ["logstd"] + [f"critic.{k}.{w}" for k in (0, 2, 4, 6) for w in ("weight", "bias")] + \
        [f"actor.{k}.{w}" for k in (0, 2, 4, 6) for w in ("weight", "bias")]
    assert sum(p.numel() for p in m.parameters()) == 177945
    assert m.state_dict()["logstd"].shape == (1, 12) and torch.all(m.state_dict()["logstd"] == -2.0)
    with pytest.raises(RuntimeError):
        m.act(torch.zeros(1, 47))   # not bound to a Learner: there is no CPU compute path
    with pytest.raises(ValueError):
        ActorCritic(10, 47, 14)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from oracle import learner as L
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d", rank=rank, world_size=world)
T, N = 4, 32                      # global batch; each rank owns N // world envs
sd = L.init_params(0, dtype=torch.float64)
buf, lo, lp = L.synthetic_rollout(T, N, seed=1, dtype=torch.float64, done_rate=0.05, timeout_rate=0.05)
omu, osig, olp = L.old_dist(sd, buf["obses"], buf["actions"])
full = L.epoch({k: v.clone() for k, v in sd.items()}, L.new_adam(sd), {k: v.clone() for k, v in buf.items()}, lo, lp, omu, osig, olp, 1e-3)
# shard: the Runner's protocol (SURVEY 8e): all-reduce (sum, sumsq, count) of the raw advantages, then the mean-gradient
n = N // world
sl = slice(rank * n, (rank + 1) * n)
sbuf = {k: v[:, sl].clone() for k, v in buf.items()}
params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
values = L.critic_value(params, sbuf["obses"], sbuf["privileged_obses"]); last_values = L.critic_value(params, lo[sl], lp[sl])
with torch.no_grad():
    sbuf["rewards"][sbuf["time_outs"]] = values[sbuf["time_outs"]]
    adv = L.gae(sbuf["rewards"], sbuf["dones"] | sbuf["time_outs"], values, last_values, 0.995, 0.95)
    returns = values + adv
    stats = torch.stack([adv.sum(), (adv ** 2).sum(), torch.tensor(float(adv.numel()), dtype=torch.float64)])
    dist.all_reduce(stats)
    mean = stats[0] / stats[2]; std = torch.sqrt((stats[1] - stats[0] * mean) / (stats[2] - 1))
    advn = (adv - mean) / (std + 1e-8)
mu = L.actor_mean(params, sbuf["obses"]); sigma = torch.exp(params["logstd"]).expand_as(mu)
logp = L.normal_log_prob(sbuf["actions"], mu, sigma).sum(-1)
loss = ((values - returns) ** 2).mean() + L.surrogate(olp[:, sl], logp, advn) + (torch.clip(mu - 1, min=0).square().mean() + torch.clip(mu + 1, max=0).square().mean()) \
    + (-0.01) * (0.5 + L.HALF_LOG_2PI + torch.log(sigma)).sum(-1).mean()
loss.backward()
worst = 0.0
for k, p in params.items():
    g = p.grad.clone(); dist.all_reduce(g); g /= world
    worst = max(worst, (g - full["grads"][k]).abs().max().item() / max(full["grads"][k].abs().max().item(), 1e-12))
assert worst < 1e-9, worst
if rank == 0: print("OK", worst)
dist.destroy_process_group()
'''


def test_sharded_gradient_protocol_equals_single_process_gloo_world2(tmp_path):
    """2 processes (gloo, CPU): per-shard mean gradients with globally all-reduced advantage moments, averaged over
    ranks, equal the single-process full-batch gradients of utils/runner.py:132-163 (fp64, 1e-9)."""
    import random

    port = random.randint(20000, 40000)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % (ROOT, port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
