"""The exact `__host__ __device__` bodies the CUDA kernels execute (csrc/t1_env.cuh, t1_dynamics.cuh, terrain.cuh), compiled
for the HOST by tests/hostcheck (g++, no FMA contraction), against the golden fixtures from the reference and the FP64
physics oracle.  Runs without a GPU; the same comparisons run on the device in tests/test_gpu_env*.py."""
import ctypes as C

import numpy as np
import pytest

from golden_util import OUT_EXACT, OUT_FLOAT, STATE_KEYS, close, fixture_cfg, load, load_cfg, step_inputs
from hostcheck_util import lib, pack_state, unpack


def _cfg_structs(cfg):
    from booster_gym_b200 import config, robot

    c = config.t1_config(cfg)
    sp = config.sim_params(cfg)
    m = robot.model_f(foot_corner=config.feet_edge_pos(cfg), dt=sp["dt"], gravity=sp["gravity"])
    return m, c


def run_host_step(z, terrain, name=None, fn="hc_env_post"):
    cfg = fixture_cfg(name, terrain)
    m, c = _cfg_structs(cfg)
    st = step_inputs(z)
    n = st["root_states"].shape[0]
    f, i, ff, fi = pack_state(st, n)
    hf = np.ascontiguousarray(z["hf"]) if "hf" in z.files else None
    table = np.ascontiguousarray(z["table"])
    obs = np.zeros((n, 47), np.float32); priv = np.zeros((n, 14), np.float32); rew = np.zeros(n, np.float32)
    done = np.zeros(n, np.uint8); touts = np.zeros(n, np.uint8); terms = np.zeros((c.n_rew, n), np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731
    getattr(lib(), fn)(C.byref(m), C.byref(c), P(hf), hf.shape[0] if hf is not None else 0, hf.shape[1] if hf is not None else 0,
                      P(f), P(i), n, P(table), C.c_longlong(int(z["common_step"])), C.c_ulonglong(1), 1, P(obs), P(priv), P(rew),
                      P(done), P(touts), P(terms))
    got = unpack(f, i, ff, fi)
    got.update(obs=obs, priv=priv, rew=rew, reset_buf=done, extras_time_outs=touts)
    return cfg, got, terms


@pytest.mark.parametrize("name,terrain", [("env_step_plane.npz", "plane"), ("env_step_trimesh.npz", "trimesh"),
                                          ("env_step_plane_noreset.npz", "plane"), ("env_step_trimesh_noreset.npz", "trimesh"),
                                          ("env_step_contacts.npz", "plane")])
def test_post_physics_body_matches_reference(name, terrain):
    from booster_gym_b200 import config

    z = load(name)
    cfg, got, terms = run_host_step(z, terrain, name)
    for k in OUT_EXACT:   # masks, counters, indices: bit-exact
        assert np.array_equal(np.asarray(got[k]).astype(np.int64).reshape(z["out_" + k].shape), z["out_" + k].astype(np.int64)), k
    for k in OUT_FLOAT:   # fp32: 1e-5 relative (north_star)
        assert close(np.asarray(got[k]).reshape(z["out_" + k].shape), z["out_" + k]), k
    names = [nm for nm, _ in config.reward_terms(cfg)]
    assert len(names) == 23
    for j, nm in enumerate(names):
        assert close(terms[j], z["out_term_" + nm]), nm


@pytest.mark.parametrize("name,terrain", [("env_step_plane.npz", "plane"), ("env_step_trimesh.npz", "trimesh"),
                                          ("env_step_plane_noreset.npz", "plane"), ("env_step_contacts.npz", "plane")])
def test_two_warp_decomposition_is_bit_identical(name, terrain):
    """k_post_pair runs one env on two warps (part A + rewards | right foot + reset + teleport + observations).  Its parts, executed on
    the host in the interleaving least like the serial body (warp 1's reset and observations BEFORE warp 0's reward terms), must leave
    exactly the bits of env_post_physics: every state row, observation, reward term and mask."""
    z = load(name)
    _, a, ta = run_host_step(z, terrain, name)
    _, b, tb = run_host_step(z, terrain, name, fn="hc_env_post_pair")
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    for k in a:
        x, y = np.ascontiguousarray(a[k]), np.ascontiguousarray(b[k])
        assert x.dtype == y.dtype and np.array_equal(x.view(np.uint8), y.view(np.uint8)), k


def test_command_curriculum_body_matches_reference():
    """the kernel bodies of the command curriculum (success counts, grid update + clamp, inverse-CDF level draw with the
    reference's transposed decoding, level -> command) against the golden produced by the reference with `curriculum: true`"""
    from golden_util import OUT_EXACT, OUT_FLOAT

    z = load("env_step_curriculum.npz")
    cfg = load_cfg("plane")
    cfg["commands"]["curriculum"] = True
    m, c = _cfg_structs(cfg)
    assert c.curriculum == 1 and c.cur_lin_levels == 10 and c.cur_success_len == 1350.0
    st = step_inputs(z)
    n = st["root_states"].shape[0]
    f, i, ff, fi = pack_state(st, n)
    prob = np.ascontiguousarray(z["in_curriculum_prob"].copy())
    table = np.ascontiguousarray(z["table"])
    obs = np.zeros((n, 47), np.float32); priv = np.zeros((n, 14), np.float32); rew = np.zeros(n, np.float32)
    done = np.zeros(n, np.uint8); tout = np.zeros(n, np.uint8); terms = np.zeros((c.n_rew, n), np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hc_env_post_curriculum(C.byref(m), C.byref(c), P(f), P(i), n, P(table), C.c_longlong(int(z["common_step"])), C.c_ulonglong(1), 1,
                                 P(prob), P(obs), P(priv), P(rew), P(done), P(tout), P(terms))
    got = unpack(f, i, ff, fi)
    assert np.array_equal(prob.view(np.uint32), z["out_curriculum_prob"].view(np.uint32))          # grid: bit-exact
    assert np.array_equal(got["env_curriculum_level"].astype(np.int64), z["out_env_curriculum_level"])
    assert np.array_equal(got["commands"].view(np.uint32), z["out_commands"].view(np.uint32))       # level -> command: bit-exact
    assert np.array_equal(done.astype(bool), z["out_reset_buf"])
    got.update(obs=obs, priv=priv, rew=rew, reset_buf=done, time_out_buf=got["time_out_buf"], extras_time_outs=tout)
    for k in OUT_EXACT:
        assert np.array_equal(np.asarray(got[k]).astype(np.int64).reshape(z["out_" + k].shape), z["out_" + k].astype(np.int64)), k
    for k in OUT_FLOAT:
        assert close(np.asarray(got[k]).reshape(z["out_" + k].shape), z["out_" + k]), k


def test_reset_body_matches_reference():
    z = load("env_reset_trimesh.npz")
    cfg = load_cfg("trimesh")
    m, c = _cfg_structs(cfg)
    st = {k: z["in_" + k].copy() for k in STATE_KEYS}
    # envs/t1.py:240-242: base-frame vectors exist from construction and are NOT refreshed by reset()
    from oracle.env_oracle import quat_rotate_inverse

    rs = st["root_states"]
    n = rs.shape[0]
    st["base_lin_vel"] = quat_rotate_inverse(rs[:, 3:7], rs[:, 7:10])
    st["base_ang_vel"] = quat_rotate_inverse(rs[:, 3:7], rs[:, 10:13])
    st["projected_gravity"] = quat_rotate_inverse(rs[:, 3:7], np.tile(np.array([0, 0, -1], np.float32), (n, 1)))
    f, i, ff, fi = pack_state(st, n)
    hf = np.ascontiguousarray(z["hf"]); table = np.ascontiguousarray(z["table"])
    obs = np.zeros((n, 47), np.float32); priv = np.zeros((n, 14), np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().hc_env_reset_all(C.byref(m), C.byref(c), P(hf), hf.shape[0], hf.shape[1], P(f), P(i), n, P(table), C.c_ulonglong(1), P(obs), P(priv))
    got = unpack(f, i, ff, fi)
    assert close(obs, z["out_obs"]) and close(priv, z["out_priv"])
    for k in ("root_states", "dof_pos", "dof_vel", "commands", "gait_frequency", "last_dof_targets", "last_root_vel"):
        assert close(got[k].reshape(z["out_" + k].shape), z["out_" + k]), k
    for k in ("episode_length_buf", "cmd_resample_time", "delay_steps"):
        assert np.array_equal(got[k].astype(np.int64), z["out_" + k]), k


def test_terrain_body_bit_exact():
    z = load("terrain_lookup.npz")
    hf = np.ascontiguousarray(z["hf"])
    L = lib()
    out = np.array([L.hc_terrain_height(hf.ctypes.data_as(C.c_void_p), hf.shape[0], hf.shape[1], 50, C.c_float(0.1), C.c_double(0.005),
                                        C.c_float(x), C.c_float(y)) for x, y in z["xy"][:, :2]], dtype=np.float32)
    assert np.array_equal(out.view(np.uint32), z["heights"].view(np.uint32))


def test_dynamics_body_matches_fp64_oracle():
    """kernel recursion (CRBA + RNE + sparse LTDL, common-reference-point spatial algebra) in double vs the independent
    dense-Jacobian oracle: agreement to 1e-10 on qacc, with and without foot contact, pushes, randomised inertias"""
    from booster_gym_b200 import robot
    from oracle import physics as op

    md = robot.model_d()
    rng = np.random.default_rng(0)
    worst = 0.0
    n_contact = 0
    for it in range(60):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        airborne = it % 2 == 0
        e = op.make_env(md, pos=(rng.normal(), rng.normal(), 2.0 if airborne else 0.64), quat=q if airborne else (0, 0, 0, 1),
                        vlin=rng.normal(size=3), wb=rng.normal(size=3) * 2, q=rng.uniform(-0.5, 0.5, 12), qd=rng.normal(size=12) * 3)
        for b in range(13):
            e.mass[b] *= rng.uniform(0.8, 1.2)
            for r in range(3):
                e.com[b][r] += rng.uniform(-0.01, 0.01)
        tau = rng.normal(size=12) * 10; pf = rng.normal(size=3) * 10; pt = rng.normal(size=3) * 2
        st, qa_o, fn_o = op.tick(md, op.Env.from_buffer_copy(e), tau, pf, pt, integrate=False)
        assert st == 0
        e2 = op.Env.from_buffer_copy(e)
        qacc = (C.c_double * 18)(); fn = (C.c_double * 2)()
        lib().hc_tick_d(C.byref(md), C.byref(e2), op._d(tau), op._d(pf), op._d(pt), None, 0, 0, 50, C.c_float(0.1), C.c_double(0.005), qacc, fn, 0)
        worst = max(worst, np.max(np.abs(qa_o - np.array(qacc))) / max(1.0, np.max(np.abs(qa_o))))
        n_contact += int(fn_o.sum() > 0)
        assert not (airborne and fn_o.sum() > 0)
        assert np.allclose(fn_o, np.array(fn), rtol=1e-9, atol=1e-9)
    assert worst < 1e-10 and n_contact >= 10


def test_physics_oracle_invariants():
    """free fall, momentum conservation in flight, energy drift bound, symmetric positive-definite M"""
    from booster_gym_b200 import robot
    from oracle import physics as op

    md = robot.model_d()
    e = op.make_env(md, pos=(0, 0, 2.0))
    st, qa, _ = op.tick(md, e, integrate=False)
    assert np.allclose(qa[:3], [0, 0, -9.81], atol=1e-12) and np.abs(qa[3:]).max() < 1e-9
    M = op.mass_matrix(md, e)
    assert np.allclose(M, M.T, atol=1e-12) and np.linalg.eigvalsh(M).min() > 0
    assert abs(M[0, 0] - 31.6144) < 1e-9 and abs(M[6 + 0, 12 + 0]) < 1e-15  # total mass; the legs do not couple
    md2 = robot.model_d(enable_contact=False, enable_limits=False, gravity=0.0)
    rng = np.random.default_rng(1)
    e = op.make_env(md2, pos=(0, 0, 5.0), vlin=rng.normal(size=3), wb=rng.normal(size=3), q=rng.uniform(-0.2, 0.2, 12) + np.array([-0.2, 0, 0, 0.6, -0.25, 0] * 2),
                    qd=rng.normal(size=12))
    E0, P0, L0 = op.energy_momentum(md2, e)
    for _ in range(200):
        op.tick(md2, e)
    E1, P1, L1 = op.energy_momentum(md2, e)
    # linear momentum (no gravity, no contact): conserved by the continuous dynamics; semi-implicit Euler in generalised
    # coordinates keeps it to O(dt) (the momentum map A(q) moves within a step) - 0.4 s at dt = 2 ms: < 0.5 %
    assert np.allclose(P0, P1, rtol=5e-3, atol=5e-2)
    c0 = L0; c1 = L1
    assert np.allclose(c0, c1, rtol=5e-3, atol=0.2)             # angular momentum about the origin, same O(dt) argument
    assert abs(E1 - E0) < 0.05 * abs(E0)                      # semi-implicit Euler energy drift over 0.4 s


def test_body_contacts_match_fp64_oracle():
    """trunk box corners and the hip-yaw / shank cylinders against the ground.  The kernel recursion (composite
    contact matrices folded into the CRBA sweep, double) vs the dense-Jacobian oracle on tumbling low robots: qacc to 1e-10,
    the net contact force per body (what contact_forces / the collision reward read) to 1e-8."""
    from booster_gym_b200 import robot
    from oracle import physics as op

    md = robot.model_d()
    rng = np.random.default_rng(3)
    worst = 0.0
    hits = np.zeros(13, int)
    for it in range(120):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        e = op.make_env(md, pos=(rng.normal(), rng.normal(), rng.uniform(0.05, 0.5)), quat=q, vlin=rng.normal(size=3), wb=rng.normal(size=3) * 2,
                        q=rng.uniform(-0.8, 0.8, 12), qd=rng.normal(size=12) * 3)
        tau = rng.normal(size=12) * 10
        st, qa_o, fn_o, bf = op.tick_f(md, op.Env.from_buffer_copy(e), tau, integrate=False)
        assert st == 0
        e2 = op.Env.from_buffer_copy(e)
        qacc = (C.c_double * 18)(); fn = (C.c_double * 2)(); co = (C.c_double * 9)()
        lib().hc_tick_d_contacts(C.byref(md), C.byref(e2), op._d(tau), op._d([0] * 3), op._d([0] * 3), None, 0, 0, 50, C.c_float(0.1),
                                 C.c_double(0.005), qacc, fn, 0, co)
        worst = max(worst, np.max(np.abs(qa_o - np.array(qacc))) / max(1.0, np.max(np.abs(qa_o))))
        f2 = (bf ** 2).sum(axis=1)
        hits += f2 > 1
        co = np.array(co)
        assert np.allclose(co[:6], f2[[3, 4, 6, 9, 10, 12]], rtol=1e-8, atol=1e-8)
        assert np.allclose(co[6:], bf[0], rtol=1e-8, atol=1e-8)
        assert not f2[[1, 2, 5, 7, 8, 11]].any()          # links without collision geometry
    assert worst < 1e-10
    assert hits[0] > 10 and min(hits[3], hits[4], hits[9], hits[10]) > 5   # every new shape was exercised


def test_fallen_robot_rests_on_its_body():
    """a robot dropped face-down / on its back / on its side comes to rest ON the trunk box and the leg cylinders (before f3 it sank
    until the soles caught it): the supporting forces add up to its weight and non-foot bodies carry most of it"""
    from booster_gym_b200 import robot
    from oracle import physics as op

    md = robot.model_d()
    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2)
    h = np.sin(np.pi / 4)
    for quat in [(0, h, 0, h), (0, -h, 0, h), (h, 0, 0, h)]:
        e = op.make_env(md, pos=(0, 0, 0.4), quat=quat, q=q0)
        for _ in range(1000):
            tau = 50 * (q0 - np.array(e.q[:])) - 1.0 * np.array(e.qd[:])
            st, _, _, bf = op.tick_f(md, e, tau)
            assert st == 0
        assert np.linalg.norm(e.vlin[:]) < 2e-2 and e.pos[2] > 0.0   # (on its side it keeps rolling slowly)
        assert abs(bf[:, 2].sum() - 31.6144 * 9.81) < 0.02 * 31.6144 * 9.81
        assert np.linalg.norm(bf[[0, 3, 4, 9, 10]], axis=1).sum() > 200.0


def test_leg_leg_contacts_match_fp64_oracle():
    """leg-leg contacts (shank / foot capsules, asset.self_collisions: 0).  Kernel recursion (double; capsule pairs
    evaluated per leg lane, implicit in the lane's own velocity) vs the dense-Jacobian oracle on legs rolled into each other:
    qacc to 1e-10, per-body net contact forces to 1e-8, and the pair forces are equal and opposite."""
    from booster_gym_b200 import robot
    from oracle import physics as op

    md = robot.model_d()
    md_off = robot.model_d(self_contact=False)
    rng = np.random.default_rng(5)
    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2)
    worst, n_self = 0.0, 0
    for it in range(150):
        airborne = it % 3 != 0
        q = q0 + rng.uniform(-0.3, 0.3, 12)
        q[1] = -rng.uniform(0.0, 0.35); q[7] = rng.uniform(0.0, 0.35)      # hip rolls inward
        q[2] = rng.uniform(-0.5, 0.5); q[8] = rng.uniform(-0.5, 0.5)
        e = op.make_env(md, pos=(0, 0, 1.5 if airborne else 0.62), vlin=rng.normal(size=3), wb=rng.normal(size=3), q=q, qd=rng.normal(size=12) * 3)
        tau = rng.normal(size=12) * 10
        st, qa_o, _, bf = op.tick_f(md, op.Env.from_buffer_copy(e), tau, integrate=False)
        _, _, _, bf_off = op.tick_f(md_off, op.Env.from_buffer_copy(e), tau, integrate=False)
        assert st == 0
        hit = np.abs(bf - bf_off).max() > 1e-9
        n_self += int(hit)
        e2 = op.Env.from_buffer_copy(e)
        qacc = (C.c_double * 18)(); fn = (C.c_double * 2)(); co = (C.c_double * 9)()
        lib().hc_tick_d_contacts(C.byref(md), C.byref(e2), op._d(tau), op._d([0] * 3), op._d([0] * 3), None, 0, 0, 50, C.c_float(0.1),
                                 C.c_double(0.005), qacc, fn, 0, co)
        worst = max(worst, np.max(np.abs(qa_o - np.array(qacc))) / max(1.0, np.max(np.abs(qa_o))))
        f2 = (bf ** 2).sum(axis=1)
        assert np.allclose(np.array(co)[:6], f2[[3, 4, 6, 9, 10, 12]], rtol=1e-8, atol=1e-8)
        if hit and airborne:
            assert np.abs(bf.sum(axis=0)).max() < 1e-9    # internal forces only
    assert worst < 1e-10 and n_self > 40


def test_legs_do_not_pass_through_each_other():
    """both hip rolls driven inwards on a floating robot: without leg-leg contacts the feet end up inside each other (centre
    distance ~0); with them the foot capsules (radius 5 cm each) keep their axes about 10 cm apart"""
    from booster_gym_b200 import robot
    from oracle import physics as op

    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2)
    tgt = q0.copy(); tgt[1] = -0.3; tgt[7] = 0.3
    kd = np.array([3, 3, 3, 3, 0.3, 0.3] * 2)
    res = {}
    for sc in (True, False):
        md = robot.model_d(self_contact=sc, gravity=0.0)
        e = op.make_env(md, pos=(0, 0, 2.0), q=q0)
        mind = 1e9
        for _ in range(1000):
            tau = np.clip(50 * (tgt - np.array(e.q[:])) - kd * np.array(e.qd[:]), -30, 30)
            st, _, _, bf = op.tick_f(md, e, tau)
            assert st == 0
            p, _ = op.feet(md, e)
            mind = min(mind, np.linalg.norm(p[0] - p[1]))
        res[sc] = (mind, np.linalg.norm(bf, axis=1))
    assert res[False][0] < 0.03 and res[True][0] > 0.075
    assert res[True][1][[6, 12]].min() > 10.0 and res[False][1].max() == 0.0


def test_bounded_sincos():
    """the 25-instruction sin / cos the physics tick uses on the device for joint angles (t1_dynamics.cuh sincos_bounded): absolute error
    below 1e-7 over [-8, 8] rad (joint limits are within +-3), including tiny arguments and multiples of pi / 4"""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-8, 8, 2_000_000), rng.uniform(-1e-3, 1e-3, 100_000), np.arange(-10, 11) * (np.pi / 4), [0.0]]).astype(np.float32)
    s = np.zeros_like(x); c = np.zeros_like(x)
    lib().hc_sincos_bounded(x.ctypes.data_as(C.c_void_p), C.c_int(len(x)), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    xd = x.astype(np.float64)
    assert np.abs(s - np.sin(xd)).max() < 1e-7 and np.abs(c - np.cos(xd)).max() < 1e-7
    assert s[-1] == 0.0 and c[-1] == 1.0


def test_fp32_tick_stays_finite_under_flailing_legs():
    """the fp32 kernel body (host build) with every contact kind switched on, driven for 1.5 s by PD targets that keep swinging the
    legs through each other and into the ground: no NaN / Inf, joint speeds and the base height stay physical (the stiff
    leg-leg and body contacts are linearly implicit; this is the long-horizon stability check the one-tick parity tests cannot give)"""
    from booster_gym_b200 import robot
    from oracle import physics as op

    mf = robot.model_f()
    rng = np.random.default_rng(11)
    q0 = np.array([-0.2, 0, 0, 0.4, -0.25, 0] * 2)
    kp = np.array([200, 200, 200, 200, 50, 50] * 2, np.float64)
    kd = np.array([5, 5, 5, 5, 1, 1] * 2, np.float64)
    lim = np.array([45, 30, 30, 60, 24, 15] * 2, np.float64)

    class EnvF(C.Structure):
        _fields_ = [("pos", C.c_float * 3), ("quat", C.c_float * 4), ("vlin", C.c_float * 3), ("wb", C.c_float * 3), ("q", C.c_float * 12),
                    ("qd", C.c_float * 12), ("mass", C.c_float * 13), ("com", C.c_float * 3 * 13), ("mu", C.c_float * 2),
                    ("kscale", C.c_float * 2), ("cscale", C.c_float * 2)]

    f32 = lambda a: (C.c_float * len(a))(*[float(v) for v in a])  # noqa: E731
    worst_qd = 0.0
    for trial in range(4):
        e = EnvF()
        e.pos[:] = [0, 0, 0.70]; e.quat[:] = [0, 0, 0, 1]
        e.q[:] = list(q0)
        for b in range(13):
            e.mass[b] = mf.mass[b]
            for r in range(3):
                e.com[b][r] = mf.ipos[b][r]
        e.mu[:] = [1.0, 1.0]; e.kscale[:] = [1.0, 1.0]; e.cscale[:] = [1.0, 1.0]
        tgt = q0.copy()
        for i in range(750):
            if i % 75 == 0:   # new targets every 0.15 s: hip rolls / yaws swing inwards and outwards, knees fold
                tgt = q0 + rng.uniform(-0.6, 0.6, 12)
                tgt[1] = rng.uniform(-0.5, 0.3); tgt[7] = rng.uniform(-0.3, 0.5)
            q = np.array(e.q[:]); qd = np.array(e.qd[:])
            tau = np.clip(kp * (tgt - q) - kd * qd, -lim, lim)
            qacc = (C.c_float * 18)(); fn = (C.c_float * 2)()
            lib().hc_tick_f(C.byref(mf), C.byref(e), f32(tau), f32([0] * 3), f32([0] * 3), None, 0, 0, 50, C.c_float(0.1), C.c_double(0.005), qacc, fn, 1)
            st = np.array(list(e.pos) + list(e.quat) + list(e.vlin) + list(e.wb) + list(e.q) + list(e.qd))
            assert np.isfinite(st).all(), (trial, i)
            worst_qd = max(worst_qd, np.abs(np.array(e.qd[:])).max())
        assert -0.05 < e.pos[2] < 1.2
    assert worst_qd < 100.0
