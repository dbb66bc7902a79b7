"""Pin the CPU oracle (oracle/) to the golden vectors produced by the REFERENCE's own code (tools/make_golden.py).
No GPU needed.  The CUDA kernels are then compared against the same fixtures and against the oracle (tests/test_gpu_*)."""
import numpy as np
import pytest
import torch

from golden_util import OUT_EXACT, OUT_FLOAT, close, fixture_cfg, load, load_cfg, model_json, step_inputs


@pytest.mark.parametrize("name,terrain", [("env_step_plane.npz", "plane"), ("env_step_trimesh.npz", "trimesh"),
                                          ("env_step_plane_noreset.npz", "plane"), ("env_step_trimesh_noreset.npz", "trimesh"),
                                          ("env_step_contacts.npz", "plane")])
def test_env_oracle_step_matches_reference(name, terrain):
    from oracle.env_oracle import EnvOracle

    z = load(name)
    cfg = fixture_cfg(name, terrain)
    hf = z["hf"] if "hf" in z.files else None
    o = EnvOracle(cfg, step_inputs(z), hf, model_json())
    out = o.step_post(z["table"], int(z["common_step"]))
    got = dict(o.s)
    got.update(obs=out["obs"], priv=out["priv"], rew=out["rew"], reset_buf=out["reset_buf"], time_out_buf=out["time_out_buf"],
               extras_time_outs=out["extras_time_outs"])
    for k in OUT_EXACT:
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), z["out_" + k].astype(np.int64)), k
    for k in OUT_FLOAT:
        assert close(np.asarray(got[k]).reshape(z["out_" + k].shape), z["out_" + k]), k
    terms = [f[len("out_term_"):] for f in z.files if f.startswith("out_term_")]
    assert len(terms) == 23
    for t in terms:
        assert close(out["terms"][t], z["out_term_" + t]), t


def test_env_oracle_command_curriculum_matches_reference():
    """SURVEY 8 f4 (envs/t1.py:391-435, `curriculum: true`): grid update on reset (success test, 4-neighbourhood, clamp), level draw
    (inverse CDF on the injected uniform; the reference's transposed index decoding), level -> command formulas, bit-exact"""
    from oracle.env_oracle import EnvOracle

    z = load("env_step_curriculum.npz")
    cfg = load_cfg("plane")
    cfg["commands"]["curriculum"] = True
    o = EnvOracle(cfg, step_inputs(z), None, model_json())
    out = o.step_post(z["table"], int(z["common_step"]))
    assert np.array_equal(o.s["curriculum_prob"].view(np.uint32), z["out_curriculum_prob"].view(np.uint32))
    assert np.array_equal(o.s["env_curriculum_level"], z["out_env_curriculum_level"])
    assert (z["out_curriculum_prob"] != z["in_curriculum_prob"]).sum() >= 10 and z["out_curriculum_prob"].max() == 1.0
    assert (z["out_env_curriculum_level"] != z["in_env_curriculum_level"]).any(axis=1).sum() >= 20
    for k in ("commands", "gait_frequency"):
        assert np.array_equal(o.s[k].view(np.uint32), z["out_" + k].view(np.uint32)), k
    for k in ("cmd_resample_time", "episode_length_buf"):
        assert np.array_equal(o.s[k], z["out_" + k]), k
    assert close(out["obs"], z["out_obs"]) and close(out["rew"], z["out_rew"])
    lev = z["out_env_curriculum_level"]
    assert float(z["out_mean_lin_vel_level"]) == pytest.approx(np.abs(lev[:, 0]).astype(np.float32).mean())
    assert int(z["out_max_ang_vel_level"]) == np.abs(lev[:, 1]).max()


def test_env_oracle_reset_matches_reference():
    from golden_util import STATE_KEYS
    from oracle.env_oracle import EnvOracle

    z = load("env_reset_trimesh.npz")
    cfg = load_cfg("trimesh")
    o = EnvOracle(cfg, {k: z["in_" + k].copy() for k in STATE_KEYS}, z["hf"], model_json())
    o.init_derived()  # envs/t1.py:240-242: base-frame vectors are computed once at construction ...
    out = o.reset_all(z["table"])
    # ... and reset() does not refresh them: obs[0:6] / priv[4:7] are stale (SURVEY 8a note 2)
    assert close(out["obs"], z["out_obs"]) and close(out["priv"], z["out_priv"])
    for k in ("root_states", "dof_pos", "dof_vel", "commands", "gait_frequency", "last_dof_targets", "last_root_vel"):
        assert close(o.s[k], z["out_" + k]), k
    for k in ("episode_length_buf", "cmd_resample_time", "delay_steps"):
        assert np.array_equal(o.s[k], z["out_" + k]), k
    # reference quirk: every env of one reset call receives the same joint offsets
    assert np.all(z["out_dof_pos"] == z["out_dof_pos"][0])


def test_terrain_oracle_bit_exact():
    from oracle.env_oracle import terrain_heights

    z = load("terrain_lookup.npz")
    h = terrain_heights(z["xy"], z["hf"])
    assert np.array_equal(h.view(np.uint32), z["heights"].view(np.uint32))
    # the C restatement used by the physics oracle agrees too
    from oracle import physics as op

    t = op.make_terrain(z["hf"])
    hc = np.array([op.terrain_height(t, float(x), float(y)) for x, y in z["xy"][:512, :2]], dtype=np.float32)
    assert np.array_equal(hc.view(np.uint32), z["heights"][:512].view(np.uint32))


def test_learner_oracle_matches_reference():
    from oracle import learner as L

    z = load("learner_small.npz")
    sd = {k[len("sd0_"):]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith("sd0_")}
    buf = {k[len("buf_"):]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith("buf_")}
    last_obs, last_priv = torch.from_numpy(z["last_obs"]), torch.from_numpy(z["last_priv"])
    omu, osig, olp = L.old_dist(sd, buf["obses"], buf["actions"])
    assert np.array_equal(olp.numpy(), z["old_logp"]) and np.array_equal(omu.numpy(), z["old_mu"])
    adam = L.new_adam(sd)
    lr = 1e-3
    for ep in range(2):
        o = L.epoch(sd, adam, buf, last_obs, last_priv, omu, osig, olp, lr)
        lr = o["lr"]
        p = f"ep{ep}_"
        # same formulas, but `x @ W.T + b` vs torch.nn.Linear's fused addmm round differently in the last bit
        assert close(o["values"].numpy(), z[p + "values"], rel=1e-6, atol=1e-7)
        assert close(o["adv_raw"].numpy(), z[p + "adv"], rel=1e-6, atol=1e-6)
        assert close(o["returns"].numpy(), z[p + "returns"], rel=1e-6, atol=1e-6)
        assert close(buf["rewards"].numpy(), z[p + "rewards"], rel=1e-6, atol=1e-7)
        assert close(o["mu"].numpy(), z[p + "mu"], rel=1e-6, atol=1e-6)
        for nm in ("value_loss", "actor_loss", "bound_loss", "entropy", "kl"):
            assert abs(o[nm] - float(z[p + nm])) <= 1e-6 * max(1.0, abs(float(z[p + nm]))), nm
        assert o["lr"] == pytest.approx(float(z[p + "lr"]), rel=1e-12)
        for k, g in o["grads"].items():
            assert close(g.numpy(), z[p + "grad_" + k], rel=1e-5, atol=1e-7), k  # reduction order differs
        for k, v in sd.items():
            # Adam's m / (sqrt(v) + eps) amplifies 1-ulp gradient differences where |g| ~ eps: compare to the step size
            assert np.abs(v.numpy() - z[p + "param_" + k]).max() <= 2e-3 * lr + 1e-7, k


def test_gae_oracle_matches_reference():
    from oracle import learner as L

    z = load("gae_small.npz")
    for c in "abcd":
        adv = L.gae(torch.from_numpy(z[c + "_r"]), torch.from_numpy(z[c + "_d"]), torch.from_numpy(z[c + "_v"]),
                    torch.from_numpy(z[c + "_lv"]), 0.995, 0.95)
        assert np.array_equal(adv.numpy(), z[c + "_adv"]), c


def test_shipped_actor_known_answer_oracle():
    """the TorchScript actor shipped with the reference evaluates actor(zeros(1,47)) to this golden vector"""
    from oracle import learner as L

    z = load("t1_actor_known_answer.npz")
    sd = L.init_params(0)
    for k in list(sd):
        if k.startswith("actor."):
            sd[k] = torch.from_numpy(z[k])
    mu = L.actor_mean(sd, torch.from_numpy(z["obs"]))
    assert np.abs(mu.numpy() - z["mu"]).max() < 1e-6
    golden0 = np.array([-0.08146450, 0.08734994, -0.01901540, 0.12740925, -0.14977062, -0.07062103, 0.03186777, -0.12124509,
                        0.09794276, -0.04422237, 0.10324907, -0.01284077])
    assert np.abs(z["mu"][0] - golden0).max() < 1e-6
