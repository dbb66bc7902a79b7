"""GPU parity of K2/K3 (post-physics step, reset) through the C-ABI: against the golden fixtures produced by the REFERENCE's own
envs/t1.py (bit-exact masks / counters / indices, 1e-5 relative floats) and, at full size, against the pinned NumPy
oracle on seeded synthetic states with identical (injected) random draws."""
import copy

import numpy as np
import pytest
import torch

from golden_util import OUT_EXACT, OUT_FLOAT, STATE_KEYS, close, fixture_cfg, load, load_cfg, model_json, step_inputs

pytestmark = pytest.mark.gpu


def make_env(cfg, n, hf=None):
    from booster_gym_b200.envs import T1

    cfg = copy.deepcopy(cfg)
    cfg["env"]["num_envs"] = n
    np.random.seed(0)
    return T1(cfg, height_field=hf)


def load_state(env, st):
    n = env.num_envs
    for name, (row, cnt) in env._ffields.items():
        if name in st:
            env._fstate[row:row + cnt].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(st[name], np.float32).reshape(n, cnt).T)))
    for name, (row, cnt) in env._ifields.items():
        if name in st:
            env._istate[row:row + cnt].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(st[name]).reshape(n, cnt).T.astype(np.int32))))


def read_state(env):
    out = {}
    f, i = env._fstate.cpu().numpy(), env._istate.cpu().numpy()
    for name, (row, cnt) in env._ffields.items():
        a = f[row:row + cnt].T.copy()
        out[name] = a[:, 0] if cnt == 1 else a
    for name, (row, cnt) in env._ifields.items():
        a = i[row:row + cnt].T.copy()
        out[name] = a[:, 0] if cnt == 1 else a
    return out


def gpu_step(cfg, st, table, common_step, hf):
    n = st["root_states"].shape[0]
    env = make_env(cfg, n, hf)
    load_state(env, st)
    env.inject_rng(torch.from_numpy(table.view(np.int32)).cuda())
    env.common_step_counter = int(common_step) - 1
    obs, rew, done, extras = env.post_physics(noise=True)
    torch.cuda.synchronize()
    got = read_state(env)
    got.update(obs=obs.cpu().numpy(), priv=extras["privileged_obs"].cpu().numpy(), rew=rew.cpu().numpy(),
               reset_buf=done.cpu().numpy(), extras_time_outs=extras["time_outs"].cpu().numpy())
    terms = {k: v.cpu().numpy() for k, v in extras["rew_terms"].items()}
    env.close()
    return got, terms


def compare(got, terms, ref, ref_terms):
    for k in OUT_EXACT:
        assert np.array_equal(np.asarray(got[k]).astype(np.int64).reshape(np.asarray(ref[k]).shape), np.asarray(ref[k]).astype(np.int64)), k
    for k in OUT_FLOAT:
        assert close(np.asarray(got[k]).reshape(np.asarray(ref[k]).shape), ref[k]), k
    assert len(ref_terms) == 23
    for nm, v in ref_terms.items():
        assert close(terms[nm], v), nm


@pytest.mark.parametrize("name,terrain", [("env_step_plane.npz", "plane"), ("env_step_trimesh.npz", "trimesh"),
                                          ("env_step_plane_noreset.npz", "plane"), ("env_step_trimesh_noreset.npz", "trimesh"),
                                          ("env_step_contacts.npz", "plane")])
def test_post_physics_kernel_matches_reference_fixture(name, terrain):
    z = load(name)
    cfg = fixture_cfg(name, terrain)
    hf = z["hf"] if "hf" in z.files else None
    got, terms = gpu_step(cfg, step_inputs(z), z["table"], z["common_step"], hf)
    ref = {k: z["out_" + k] for k in OUT_EXACT + OUT_FLOAT}
    ref_terms = {f[len("out_term_"):]: z[f] for f in z.files if f.startswith("out_term_")}
    compare(got, terms, ref, ref_terms)


def test_reset_kernel_matches_reference_fixture():
    from oracle.env_oracle import quat_rotate_inverse

    z = load("env_reset_trimesh.npz")
    cfg = load_cfg("trimesh")
    st = {k: z["in_" + k].copy() for k in STATE_KEYS}
    rs = st["root_states"]
    n = rs.shape[0]
    st["base_lin_vel"] = quat_rotate_inverse(rs[:, 3:7], rs[:, 7:10])
    st["base_ang_vel"] = quat_rotate_inverse(rs[:, 3:7], rs[:, 10:13])
    st["projected_gravity"] = quat_rotate_inverse(rs[:, 3:7], np.tile(np.array([0, 0, -1], np.float32), (n, 1)))
    env = make_env(cfg, n, z["hf"])
    load_state(env, st)
    env.inject_rng(torch.from_numpy(z["table"].view(np.int32)).cuda())
    obs, extras = env.reset()
    torch.cuda.synchronize()
    got = read_state(env)
    assert close(obs.cpu().numpy(), z["out_obs"]) and close(extras["privileged_obs"].cpu().numpy(), z["out_priv"])
    for k in ("root_states", "dof_pos", "dof_vel", "commands", "gait_frequency", "last_dof_targets", "last_root_vel"):
        assert close(got[k].reshape(z["out_" + k].shape), z["out_" + k]), k
    for k in ("episode_length_buf", "cmd_resample_time", "delay_steps"):
        assert np.array_equal(got[k].astype(np.int64), z["out_" + k]), k


@pytest.mark.parametrize("terrain,common_step", [("plane", 500), ("trimesh", 500), ("trimesh", 50), ("plane", 7)])
def test_post_physics_kernel_matches_oracle_full_size(terrain, common_step):
    """N = 4096 (BASELINE configs[1]) / trimesh: the kernel against the pinned NumPy oracle, identical injected draws"""
    from oracle.env_oracle import EnvOracle
    from oracle.ref_harness import make_table
    from oracle.synth import synthetic_state

    n = 4096
    cfg = load_cfg(terrain)
    hf = load("terrain_lookup.npz")["hf"] if terrain == "trimesh" else None
    st = synthetic_state(n, 1000 + common_step, terrain == "trimesh")
    g = np.random.default_rng(3)
    st["actions"] = g.uniform(-1, 1, (n, 12)).astype(np.float32)
    st["torques"] = g.normal(0, 8, (n, 12)).astype(np.float32)
    # the 'still' draw: the oracle, like the kernel, uses the per-env Bernoulli form (u < still_proportion)
    table = make_table(n, 5)
    o = EnvOracle(cfg, st, hf, model_json())
    out = o.step_post(table, common_step)
    ref = dict(o.s)
    ref.update(obs=out["obs"], priv=out["priv"], rew=out["rew"], reset_buf=out["reset_buf"], time_out_buf=out["time_out_buf"],
               extras_time_outs=out["extras_time_outs"])
    got, terms = gpu_step(cfg, st, table, common_step, hf)
    compare(got, terms, ref, out["terms"])
    assert out["reset_buf"].sum() > 100 and out["time_out_buf"].sum() > 100


def _curriculum_cfg():
    cfg = load_cfg("plane")
    cfg["commands"]["curriculum"] = True
    return cfg


def test_command_curriculum_kernels_match_reference_fixture():
    """SURVEY 8 f4 (envs/t1.py:391-435): k_post<1> -> k_curriculum_apply -> k_post<2> through the public T1 API against the golden
    the reference produced with `curriculum: true`: grid bit-exact, levels exact, commands bit-exact, everything else as usual"""
    z = load("env_step_curriculum.npz")
    cfg = _curriculum_cfg()
    st = step_inputs(z)
    n = st["root_states"].shape[0]
    env = make_env(cfg, n)
    load_state(env, st)
    env.curriculum_prob = torch.from_numpy(z["in_curriculum_prob"])          # the checkpoint-loading path (utils/runner.py:91)
    env.inject_rng(torch.from_numpy(z["table"].view(np.int32)).cuda())
    env.common_step_counter = int(z["common_step"]) - 1
    obs, rew, done, extras = env.post_physics(noise=True)
    torch.cuda.synchronize()
    got = read_state(env)
    assert np.array_equal(env.curriculum_prob.cpu().numpy().view(np.uint32), z["out_curriculum_prob"].view(np.uint32))
    assert np.array_equal(got["env_curriculum_level"].astype(np.int64), z["out_env_curriculum_level"])
    assert np.array_equal(got["commands"].view(np.uint32), z["out_commands"].view(np.uint32))
    assert float(env.mean_lin_vel_level) == pytest.approx(float(z["out_mean_lin_vel_level"]))
    assert float(env.mean_ang_vel_level) == pytest.approx(float(z["out_mean_ang_vel_level"]))
    assert int(env.max_lin_vel_level) == int(z["out_max_lin_vel_level"]) and int(env.max_ang_vel_level) == int(z["out_max_ang_vel_level"])
    got.update(obs=obs.cpu().numpy(), priv=extras["privileged_obs"].cpu().numpy(), rew=rew.cpu().numpy(),
               reset_buf=done.cpu().numpy(), extras_time_outs=extras["time_outs"].cpu().numpy())
    terms = {k: v.cpu().numpy() for k, v in extras["rew_terms"].items()}
    ref = {k: z["out_" + k] for k in OUT_EXACT + OUT_FLOAT}
    ref_terms = {f[len("out_term_"):]: z[f] for f in z.files if f.startswith("out_term_")}
    compare(got, terms, ref, ref_terms)
    env.close()


def test_command_curriculum_full_size_and_closed_loop():
    """N = 4096: kernel vs the pinned oracle on a synthetic state with a partly filled grid (identical injected draws), then 300 real
    steps of the public API with the curriculum on: the grid only grows, stays in [0, 1], levels stay inside the grid"""
    from oracle.env_oracle import EnvOracle
    from oracle.ref_harness import make_table
    from oracle.synth import synthetic_state

    n = 4096
    cfg = _curriculum_cfg()
    L, A = cfg["commands"]["lin_vel_levels"], cfg["commands"]["ang_vel_levels"]
    st = synthetic_state(n, 77, False)
    g = np.random.default_rng(9)
    st["actions"] = g.uniform(-1, 1, (n, 12)).astype(np.float32)
    st["torques"] = g.normal(0, 8, (n, 12)).astype(np.float32)
    st["root_states"][:, 7:13] *= 0.2
    to = np.arange(1, n, 9)                           # the envs whose episode times out this step: make them track their command
    st["filtered_lin_vel"][to, 0:2] = st["commands"][to, 0:2] / np.float32(0.9)
    st["filtered_ang_vel"][to, 2] = st["commands"][to, 2] / np.float32(0.9)
    prob = np.zeros((2 * L + 1, 2 * A + 1), np.float32)
    prob[L - 4: L + 5, A - 4: A + 5] = g.uniform(0, 1, (9, 9)).astype(np.float32)
    st["curriculum_prob"] = prob
    st["env_curriculum_level"] = np.stack([g.integers(-L, L + 1, n), g.integers(-A, A + 1, n)], axis=1).astype(np.int64)
    table = make_table(n, 6)
    o = EnvOracle(cfg, st, None, model_json())
    out = o.step_post(table, 500)
    env = make_env(cfg, n)
    load_state(env, st)
    env.curriculum_prob = torch.from_numpy(prob)
    env.inject_rng(torch.from_numpy(table.view(np.int32)).cuda())
    env.common_step_counter = 499
    obs, rew, done, extras = env.post_physics(noise=True)
    torch.cuda.synchronize()
    got = read_state(env)
    assert (o.s["curriculum_prob"] != prob).sum() > 50
    assert np.array_equal(env.curriculum_prob.cpu().numpy().view(np.uint32), o.s["curriculum_prob"].view(np.uint32))
    assert np.array_equal(got["env_curriculum_level"].astype(np.int64), o.s["env_curriculum_level"])
    assert np.array_equal(got["commands"].view(np.uint32), o.s["commands"].view(np.uint32))
    assert close(obs.cpu().numpy(), out["obs"]) and close(rew.cpu().numpy(), out["rew"])
    env.close()
    # closed loop
    env = make_env(cfg, 1024)
    obs, _ = env.reset()
    before = env.curriculum_prob.clone()
    act = torch.zeros(1024, 12, device="cuda")
    for _ in range(300):
        obs, rew, done, extras = env.step(act)
    torch.cuda.synchronize()
    p = env.curriculum_prob
    assert torch.isfinite(obs).all() and p.min() >= 0 and p.max() <= 1.0 and (p >= before).all()
    lev = env.env_curriculum_level
    assert lev[:, 0].abs().max() <= L and lev[:, 1].abs().max() <= A
    env.close()
