"""GPU tests that close SURVEY.md section 8 rows the round-1 review found untested:

* a21  one-off domain randomisation of `_create_envs` (envs/t1.py:69-83,139-167, ranges envs/T1.yaml:208-249)
* a10  the decimated PD law against the arrays the REFERENCE's own loop left behind (`post_loop_*` of every
       tests/golden/env_step_*.npz; envs/t1.py:443-456)
* f1   device-side episode statistics against the reference Recorder's per-step bookkeeping (utils/recorder.py:36-62)
"""
import copy

import numpy as np
import pytest
import torch

from golden_util import STATE_KEYS, fixture_cfg, load

pytestmark = pytest.mark.gpu


def _make(cfg, n, seed=42):
    from booster_gym_b200.envs import T1

    cfg = copy.deepcopy(cfg)
    cfg["env"]["num_envs"] = n
    cfg["basic"]["seed"] = seed
    cfg["basic"]["headless"] = True
    np.random.seed(seed)
    return T1(cfg)


def _uniform_ok(x, lo, hi, n):
    """x ~ U(lo, hi): support, mean and variance within 5 standard errors"""
    x = np.asarray(x, np.float64).ravel()
    w = hi - lo
    assert x.min() >= lo - 1e-6 * max(1.0, abs(lo)) and x.max() <= hi + 1e-6 * max(1.0, abs(hi)), (x.min(), x.max(), lo, hi)
    assert abs(x.mean() - 0.5 * (lo + hi)) < 5.0 * w / np.sqrt(12.0 * x.size), (x.mean(), lo, hi)
    assert abs(x.var() - w * w / 12.0) < 5.0 * (w * w / 12.0) * np.sqrt(0.8 / x.size) + 1e-12, (x.var(), w * w / 12.0)
    # coverage: no empty decile
    hist, _ = np.histogram(x, bins=10, range=(lo, hi))
    assert hist.min() > 0.7 * x.size / 10, hist


def _max_offdiag_corr(a):
    a = np.asarray(a, np.float64)
    c = np.corrcoef(a, rowvar=False)
    return np.abs(c - np.eye(c.shape[0])).max()


def test_dr_init_matches_the_reference_distributions(t1_cfg):
    """k_init_params vs envs/t1.py:69-83 (PD gains, joint friction: one sample per env x DoF), :139-159 (Trunk CoM +-0.1 m and
    mass x U(.8,1.2), other bodies CoM +-5 mm and mass x U(.98,1.02)), :162-167 (foot material), envs/T1.yaml:208-249.
    `base_mass_scaled` holds the RAW U[0,1) samples (utils/utils.py:27-28, SURVEY 8a note 5)."""
    from booster_gym_b200 import robot

    n = 4096
    env = _make(t1_cfg, n)
    md = robot.model_d(foot_corner=[list(map(float, r)) for r in t1_cfg["asset"]["feet_edge_pos"]])
    kp_nom = torch.tensor(list(env._c_cfg.kp_nominal)).numpy()
    kd_nom = torch.tensor(list(env._c_cfg.kd_nominal)).numpy()
    assert list(kp_nom) == [200, 200, 200, 200, 50, 50] * 2 and list(kd_nom) == [5, 5, 5, 5, 1, 1] * 2   # envs/T1.yaml:92-93
    kp = env.dof_stiffness.cpu().numpy() / kp_nom
    kd = env.dof_damping.cpu().numpy() / kd_nom
    fr = env.dof_friction.cpu().numpy()
    for j in range(12):
        _uniform_ok(kp[:, j], 0.95, 1.05, n)
        _uniform_ok(kd[:, j], 0.95, 1.05, n)
        _uniform_ok(fr[:, j], 0.0, 2.0, n)
    # one independent draw per env x DoF x quantity: no correlation across DoFs, across quantities, or between neighbouring envs
    lim = 5.0 / np.sqrt(n)
    assert _max_offdiag_corr(np.concatenate([kp, kd, fr], axis=1)) < lim
    assert abs(np.corrcoef(kp[:-1, 0], kp[1:, 0])[0, 1]) < lim and abs(np.corrcoef(fr[:-1, 3], fr[1:, 3])[0, 1]) < lim

    mass = env._fview("body_mass").cpu().numpy()          # [N, 13]
    com = env._fview("body_com").cpu().numpy().reshape(n, 13, 3)
    m0 = np.array([md.mass[b] for b in range(13)])
    c0 = np.array([[md.ipos[b][k] for k in range(3)] for b in range(13)])
    assert abs(m0.sum() - 31.6144) < 1e-3
    _uniform_ok(mass[:, 0] / m0[0], 0.8, 1.2, n)
    for k in range(3):
        _uniform_ok(com[:, 0, k] - c0[0, k], -0.1, 0.1, n)
    for b in range(1, 13):
        _uniform_ok(mass[:, b] / m0[b], 0.98, 1.02, n)
        for k in range(3):
            _uniform_ok(com[:, b, k] - c0[b, k], -0.005, 0.005, n)
    allp = np.concatenate([mass / m0, (com - c0).reshape(n, 39)], axis=1)
    assert _max_offdiag_corr(allp) < lim

    # raw samples, not the applied offsets (quirk 5): the applied values are lo + (hi - lo) * raw
    raw = env.base_mass_scaled.cpu().numpy()
    for k in range(4):
        _uniform_ok(raw[:, k], 0.0, 1.0, n)
    assert np.abs((com[:, 0, :] - c0[0]) - (-0.1 + 0.2 * raw[:, 0:3])).max() < 1e-6
    assert np.abs(mass[:, 0] / m0[0] - (0.8 + 0.4 * raw[:, 3])).max() < 1e-6

    # foot material (envs/t1.py:162-167): friction U(.1, 2) averaged with the ground's (PhysX default combine mode),
    # compliance U(.5, 1.5) and restitution U(.1, .9) enter this build's contact law as 1 / compliance and 1 - restitution / 2
    mu_g = float(t1_cfg["terrain"]["static_friction"]) if "static_friction" in t1_cfg["terrain"] else float(env._c_cfg.terrain_friction)
    mu = 2.0 * env._fview("foot_friction").cpu().numpy() - mu_g
    comp = 1.0 / env._fview("foot_kscale").cpu().numpy()
    rest = 2.0 * (1.0 - env._fview("foot_cscale").cpu().numpy())
    for k in range(2):
        _uniform_ok(mu[:, k], 0.1, 2.0, n)
        _uniform_ok(comp[:, k], 0.5, 1.5, n)
        _uniform_ok(rest[:, k], 0.1, 0.9, n)
    assert _max_offdiag_corr(np.concatenate([mu, comp, rest], axis=1)) < lim


def test_dr_init_disabled_and_gaussian(t1_cfg):
    """apply_randomization(x, None) is the identity (utils/utils.py:6-7); gaussian uses range = [mean, "var" used as std] (:19-21)"""
    n = 4096
    cfg = copy.deepcopy(t1_cfg)
    for k in ("dof_stiffness", "dof_damping", "base_com", "base_mass", "other_com", "other_mass", "friction", "compliance", "restitution"):
        cfg["randomization"][k] = None
    cfg["randomization"]["dof_friction"] = {"range": [0.5, 0.1], "operation": "additive", "distribution": "gaussian"}
    env = _make(cfg, n)
    assert torch.equal(env.dof_stiffness, torch.tensor(list(env._c_cfg.kp_nominal), device="cuda").expand(n, 12))
    assert torch.equal(env.dof_damping, torch.tensor(list(env._c_cfg.kd_nominal), device="cuda").expand(n, 12))
    assert (env.base_mass_scaled == 0).all()
    m = env._fview("body_mass")
    assert (m == m[0:1]).all()
    fr = env.dof_friction.cpu().numpy().astype(np.float64)
    assert abs(fr.mean() - 0.5) < 5 * 0.1 / np.sqrt(fr.size) and abs(fr.std() - 0.1) < 0.002
    assert _max_offdiag_corr(fr) < 5.0 / np.sqrt(n)


@pytest.mark.parametrize("name,terrain", [("env_step_plane.npz", "plane"), ("env_step_trimesh.npz", "trimesh"),
                                          ("env_step_plane_noreset.npz", "plane"), ("env_step_contacts.npz", "plane")])
def test_pd_law_matches_the_reference_loop(name, terrain):
    """a10: the fixtures were produced by the reference's real `T1.step()` with `gym.simulate` stubbed to the identity, so its
    decimation loop (envs/t1.py:443-456) ran on a frozen (q, qd) and left `actions` (clipped), `last_dof_targets` (switched at
    substep == delay_steps) and the MEAN torque behind.  The kernel's PD code is driven the same way: ten single-substep launches
    with the state restored in between, the delay armed only in the launch whose index equals the env's delay."""
    from test_gpu_env_post import load_state, make_env

    z = load(name)
    cfg = fixture_cfg(name, terrain)
    st = {k: z["in_" + k].copy() for k in STATE_KEYS}
    n = st["dof_pos"].shape[0]
    env = make_env(cfg, n, z["hf"] if "hf" in z.files else None)
    load_state(env, st)
    raw = torch.from_numpy(z["actions_raw"]).cuda()
    frozen = env._fstate.clone()
    delay = torch.from_numpy(st["delay_steps"]).cuda()
    tsum = torch.zeros(n, 12, device="cuda")
    last = env.last_dof_targets.clone()
    for i in range(10):
        env._fstate.copy_(frozen)
        env.last_dof_targets.copy_(last)
        env.delay_steps.copy_(torch.where(delay == i, torch.zeros_like(delay), torch.full_like(delay, -1)).to(env.delay_steps.dtype))
        env.physics(raw, 1, apply_pd=True)
        tsum += env.torques
        last = env.last_dof_targets.clone()
        acts = env.actions.clone()
    torch.cuda.synchronize()
    assert np.array_equal(acts.cpu().numpy(), z["post_loop_actions"])                    # torch.clip: exact
    assert np.array_equal(last.cpu().numpy(), z["post_loop_last_dof_targets"])           # one fp32 multiply-add: exact
    got, ref = (tsum / 10.0).cpu().numpy(), z["post_loop_torques"]
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    assert np.abs(ref).max() > 1.0   # the fixture exercises the law (non-trivial torques, some at the clip limits)


def _reference_recorder_means(dones, rews, terms):
    """utils/recorder.py:36-62 restated on host tensors: per-step bookkeeping of the reference Recorder, returning the means it
    would write (`steps`, `reward`, `episode/<term>`) and the number of finished episodes"""
    episode_steps = None
    stats = {}
    last = {"steps": []}
    for t in range(len(dones)):
        done = dones[t]
        if episode_steps is None:
            episode_steps = torch.zeros_like(done, dtype=torch.int64)        # :37-38 (the first call does not count a step)
        else:
            episode_steps += 1                                                # :39-40
        last["steps"].extend(episode_steps[done].tolist())                   # :41-42
        episode_steps[done] = 0                                               # :43
        ep_info = {"reward": rews[t]}
        ep_info.update({k: v[t] for k, v in terms.items()})                  # utils/runner.py:119-120
        for key, value in ep_info.items():
            if key not in stats:
                stats[key] = torch.zeros_like(value)                         # :46-47
            stats[key] += value                                              # :48
            last.setdefault(key, []).extend(stats[key][done].tolist())       # :49-52
            stats[key][done] = 0                                             # :53
    cnt = len(last["steps"])
    return {k: (sum(v) / len(v) if len(v) else 0.0) for k, v in last.items()}, cnt


def test_episode_stats_match_the_reference_recorder(t1_cfg):
    """f1: b200_t1_episode_stats (sums accumulated inside k_post) against the reference Recorder fed with the same rollout's
    per-step `done`, `rew` and `infos["rew_terms"]` (utils/runner.py:119-121): same episode count, `reward` and every
    `episode/<term>` mean to 1e-5.  `steps`: the reference's counter is 0 (not 1) after its very first call, so the FIRST episode of
    every env is reported one step short and all later ones exactly; the device counter starts at -1 to report the same numbers."""
    n = 512
    cfg = copy.deepcopy(t1_cfg)
    cfg["terrain"]["type"] = "trimesh"
    env = _make(cfg, n)
    env.reset()
    env.episode_stats()   # clear
    g = torch.Generator(device="cuda").manual_seed(1)
    dones, rews, terms = [], [], {k: [] for k in env.reward_names}
    for s in range(160):
        obs, rew, done, extras = env.step(torch.randn(n, 12, device="cuda", generator=g) * 0.7)
        dones.append(done.clone().cpu())
        rews.append(rew.clone().cpu())
        for k in env.reward_names:
            terms[k].append(extras["rew_terms"][k].clone().cpu())
    ref, cnt_ref = _reference_recorder_means(dones, rews, terms)
    got, cnt = env.episode_stats()
    assert cnt == cnt_ref and cnt > 50
    for k in ["reward"] + list(env.reward_names):
        assert abs(got[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
    assert abs(got["steps"] - ref["steps"]) <= 1e-9 * max(1.0, ref["steps"]), (got["steps"], ref["steps"])
