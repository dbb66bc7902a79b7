#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code (read-only /root/reference).

    python tools/make_golden.py            # needs the reference tree; run in the build container

Fixtures (small .npz files, committed; the GPU box has no reference tree):
  env_step_{plane,trimesh}.npz  reference T1.step() with physics = identity on a seeded synthetic state, RNG injected
  env_step_curriculum.npz       reference T1.step() with `commands.curriculum: true`
  env_step_contacts.npz         reference T1.step() with non-zero contact_forces and `terminate_contacts_on: [Trunk, Shank]`
                                (collision reward, contact termination; `--contacts-only` regenerates just this one)
  env_reset_trimesh.npz         reference T1.reset()
  terrain_lookup.npz            reference Terrain.terrain_heights on a seeded heightfield
  learner_small.npz             reference ActorCritic + discount_values + surrogate_loss + the loop body of
                                utils/runner.py:123-180 (two epochs, torch.optim.Adam, clip_grad_norm_)
  gae_small.npz                 reference discount_values alone
  t1_actor_known_answer.npz     the shipped TorchScript actor deploy/models/T1.pt: weights + input/output pairs
The oracle (oracle/*.py, oracle/physics_oracle.c) and the CUDA kernels are both tested against these files.
"""
import copy
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_cfg(terrain):
    cfg = yaml.load(open(os.path.join(ROOT, "envs", "T1.yaml")).read(), Loader=yaml.FullLoader)
    cfg = copy.deepcopy(cfg)
    cfg["terrain"]["type"] = terrain
    cfg["viewer"]["record_video"] = False
    cfg["basic"]["headless"] = True
    return cfg


from oracle.synth import synthetic_state  # noqa: E402


def run_reference_step(mods, cfg, state, hf, common_step_before, actions, table_seed):
    n = state["root_states"].shape[0]
    e = H.build_reference_env(mods, cfg, state, hf)
    e.common_step_counter = common_step_before
    # which envs will resample a command this step is decided inside step(); the 'still' set must have exactly
    # int(still_proportion * len) members (randperm semantics) -> pre-compute it with a dry run without randomness use
    dry = H.build_reference_env(mods, cfg, state, hf)
    dry.common_step_counter = common_step_before
    inj0 = H.Injector(H.make_table(n, table_seed), n)
    H.instrument(dry, inj0, cfg)
    inj0.randperm = lambda m, **kw: torch.arange(m)
    with H.patched_rng(inj0):
        torch.randperm = inj0.randperm
        dry.step(torch.from_numpy(actions))
    # envs that resampled = those whose cmd_resample_time changed (or reset to 0 then advanced)
    resampled = (dry.cmd_resample_time.numpy() != state["cmd_resample_time"]) | \
                ((dry.episode_length_buf.numpy() == 0) & (dry.cmd_resample_time.numpy() > 0))
    ids = np.nonzero(resampled)[0]
    still = np.zeros(n, bool)
    still[ids[: int(cfg["commands"]["still_proportion"] * len(ids))]] = True
    table = H.make_table(n, table_seed, still_envs=still, still_proportion=cfg["commands"]["still_proportion"])
    inj = H.Injector(table, n)
    cap = H.instrument(e, inj, cfg)
    with H.patched_rng(inj):
        e.step(torch.from_numpy(actions))
    assert np.array_equal((e.cmd_resample_time.numpy() != state["cmd_resample_time"]) |
                          ((e.episode_length_buf.numpy() == 0) & (e.cmd_resample_time.numpy() > 0)), resampled)
    out = H.snapshot(e)
    return table, cap, out, inj.log


def gen_env(mods):
    rngs = {"plane": 11, "trimesh": 12}
    for terrain in ("plane", "trimesh"):
        cfg = load_cfg(terrain)
        hf = None
        if terrain == "trimesh":
            np.random.seed(42)
            t = mods[1].Terrain.__new__(mods[1].Terrain)
            t.terrain_cfg = cfg["terrain"]
            t.gym, t.sim, t.device, t.type = H.NullGym({}), None, "cpu", "trimesh"
            t._create_trimesh()
            hf = t.height_field_raw.copy()
        n = 96
        state = synthetic_state(n, rngs[terrain], terrain == "trimesh")
        actions = np.random.default_rng(5).uniform(-1.4, 1.4, (n, 12)).astype(np.float32)
        # common_step_counter becomes 500 inside step(): kick (100 | 500) and push (250 | 500) both fire
        table, cap, out, log = run_reference_step(mods, cfg, state, hf, 499, actions, 77)
        save = {"in_" + k: v for k, v in state.items()}
        save.update({"out_" + k: v for k, v in out.items()})
        save.update(actions_raw=actions, table=table, common_step=np.int64(500), post_loop_torques=cap["torques"],
                    post_loop_last_dof_targets=cap["last_dof_targets"], post_loop_actions=cap["actions"])
        if hf is not None:
            save["hf"] = hf
        np.savez_compressed(os.path.join(OUT, f"env_step_{terrain}.npz"), **save)
        print(terrain, "step: resets", int(out["reset_buf"].sum()), "time_outs", int(out["time_out_buf"].sum()),
              "contacts", int(out["feet_contact"].sum()), "rng calls", len(log))
        # a second case: push phase == push_duration (forces zeroed), no kick, nobody resets -> extras["time_outs"] stays stale
        st2 = synthetic_state(n, rngs[terrain] + 100, terrain == "trimesh", ep_case=False)
        st2["root_states"][:, 2] = 0.7 + (0.0 if hf is None else 0.2)
        st2["root_states"][:, 7:13] *= 0.3
        st2["root_states"][: n // 4, 0] = np.random.default_rng(1).uniform(10, 60, n // 4)
        st2["root_states"][: n // 4, 1] = np.random.default_rng(2).uniform(2, 8, n // 4)
        table, cap, out, log = run_reference_step(mods, cfg, st2, hf, 299, actions, 78)
        assert out["reset_buf"].sum() == 0
        save = {"in_" + k: v for k, v in st2.items()}
        save.update({"out_" + k: v for k, v in out.items()})
        save.update(actions_raw=actions, table=table, common_step=np.int64(300), post_loop_torques=cap["torques"],
                    post_loop_last_dof_targets=cap["last_dof_targets"], post_loop_actions=cap["actions"])
        if hf is not None:
            save["hf"] = hf
        np.savez_compressed(os.path.join(OUT, f"env_step_{terrain}_noreset.npz"), **save)
        print(terrain, "no-reset step: time_outs", int(out["time_out_buf"].sum()), "extras", int(out["extras_time_outs"].sum()))
        if terrain == "plane":
            # command curriculum (envs/t1.py:391-435, `curriculum: true`): successful episodes (time-out with the command tracked)
            # raise the grid, failures do not; resampling envs draw a cell and derive their command from its levels
            cfgc = copy.deepcopy(cfg)
            cfgc["commands"]["curriculum"] = True
            stc = synthetic_state(n, 41, False)
            gg = np.random.default_rng(7)
            L, A = cfgc["commands"]["lin_vel_levels"], cfgc["commands"]["ang_vel_levels"]
            prob = np.zeros((2 * L + 1, 2 * A + 1), np.float32)
            prob[L - 2: L + 3, A - 3: A + 4] = gg.uniform(0.0, 1.0, (5, 7)).astype(np.float32)
            prob[L, A] = 1.0
            prob[L + 1, A + 1] = 0.95          # += 0.1 must clamp to 1
            lev = np.stack([gg.integers(-2, 3, n), gg.integers(-3, 4, n)], axis=1).astype(np.int64)
            lev[0] = (-L, A)                   # grid corner: two neighbours fall outside
            lev[9] = (1, 1)
            stc["curriculum_prob"], stc["env_curriculum_level"] = prob, lev
            # envs 1, 10, 19, ... time out this step (ep_len 1500 -> 1501): make most of them successes, a few near misses
            stc["root_states"][:, 2] = 0.75
            stc["root_states"][:, 7:13] *= 0.2
            to = np.arange(1, n, 9)
            stc["commands"][to] = gg.uniform(-0.5, 0.5, (len(to), 3)).astype(np.float32)
            # the step first filters: filtered = 0.1 * base_vel + 0.9 * filtered -> choose filtered so that the error is small
            stc["filtered_lin_vel"][to, 0:2] = stc["commands"][to, 0:2] / np.float32(0.9)
            stc["filtered_ang_vel"][to, 2] = stc["commands"][to, 2] / np.float32(0.9)
            stc["filtered_lin_vel"][to[2], 0] += 0.9      # x error > 0.4: not a success
            stc["filtered_ang_vel"][to[3], 2] -= 0.5      # yaw error > 0.2: not a success
            stc["episode_length_buf"][to[4]] = 1340        # early termination below the length tolerance (fails by height below)
            stc["root_states"][to[4], 2] = 0.3
            stc["episode_length_buf"][to[5]] = 1360        # terminated by height but long enough and on target: a success
            stc["root_states"][to[5], 2] = 0.3
            table, cap, out, log = run_reference_step(mods, cfgc, stc, None, 499, actions, 81)
            save = {"in_" + k: v for k, v in stc.items()}
            save.update({"out_" + k: v for k, v in out.items()})
            save.update(actions_raw=actions, table=table, common_step=np.int64(500), post_loop_torques=cap["torques"],
                        post_loop_last_dof_targets=cap["last_dof_targets"], post_loop_actions=cap["actions"])
            np.savez_compressed(os.path.join(OUT, "env_step_curriculum.npz"), **save)
            print("curriculum step: resets", int(out["reset_buf"].sum()), "grid delta", float((out["curriculum_prob"] - prob).sum()),
                  "cells raised", int((out["curriculum_prob"] != prob).sum()), "levels changed", int((out["env_curriculum_level"] != lev).any(axis=1).sum()))
        if terrain == "trimesh":
            # reset(): every env resets, commands resampled for all, observations
            state3 = synthetic_state(n, 33, True)
            e = H.build_reference_env(mods, cfg, state3, hf)
            still = np.zeros(n, bool); still[: int(0.1 * n)] = True
            table = H.make_table(n, 79, still_envs=still)
            inj = H.Injector(table, n)
            H.instrument(e, inj, cfg)
            with H.patched_rng(inj):
                e.reset()
            out = H.snapshot(e)
            save = {"in_" + k: v for k, v in state3.items()}
            save.update({"out_" + k: v for k, v in out.items()})
            save.update(table=table, hf=hf)
            np.savez_compressed(os.path.join(OUT, "env_reset_trimesh.npz"), **save)
            # terrain lookup alone
            gg = np.random.default_rng(0)
            K = 4096
            xy = np.stack([gg.uniform(-4.9, 84.7, K), gg.uniform(-4.9, 14.7, K), gg.uniform(0, 1, K)], axis=1).astype(np.float32)
            xy[:256, 0:2] = (np.round(xy[:256, 0:2] * 10) / 10).astype(np.float32)
            xy[256:320, 0] = gg.uniform(-5.09, -5.0, 64).astype(np.float32)
            h = e.terrain.terrain_heights(torch.from_numpy(xy)).numpy()
            np.savez_compressed(os.path.join(OUT, "terrain_lookup.npz"), hf=hf, xy=xy, heights=h)


def gen_contacts(mods):
    """net contact forces on non-foot bodies -> `collision` reward (envs/t1.py:627-629) and contact termination
    (:553, with terminate_contacts_on set).  contact_forces is an INPUT of the reference step (gym tensor, physics = identity)."""
    cfg = load_cfg("plane")
    cfg["rewards"]["terminate_contacts_on"] = ["Trunk", "Shank"]
    n = 96
    state = synthetic_state(n, 11, False)
    state["root_states"][:, 2] = 0.7            # nobody terminates by height: resets below come from contacts / velocity / length
    gg = np.random.default_rng(21)
    cf = np.zeros((n, 13, 3), np.float32)
    hit = gg.uniform(size=(n, 13)) < 0.25
    mag = np.where(gg.uniform(size=(n, 13)) < 0.5, gg.uniform(0.2, 0.999, (n, 13)), gg.uniform(1.001, 300.0, (n, 13)))
    d = gg.normal(size=(n, 13, 3))
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    cf[hit] = (d * mag[:, :, None])[hit].astype(np.float32)
    cf[:, [1, 2, 5, 7, 8, 11]] = 0.0             # links without collision geometry never carry contact forces
    state["contact_forces"] = cf
    actions = np.random.default_rng(5).uniform(-1.4, 1.4, (n, 12)).astype(np.float32)
    table, cap, out, log = run_reference_step(mods, cfg, state, None, 499, actions, 83)
    save = {"in_" + k: v for k, v in state.items()}
    save.update({"out_" + k: v for k, v in out.items()})
    save.update(actions_raw=actions, table=table, common_step=np.int64(500), post_loop_torques=cap["torques"],
                post_loop_last_dof_targets=cap["last_dof_targets"], post_loop_actions=cap["actions"])
    np.savez_compressed(os.path.join(OUT, "env_step_contacts.npz"), **save)
    print("contacts step: resets", int(out["reset_buf"].sum()), "collision term min", float(out["term_collision"].min()),
          "envs with collisions", int((out["term_collision"] != 0).sum()))


def gen_learner(mods):
    import torch.nn.functional as F

    _, _, U, M = mods
    sys.path.insert(0, ROOT)
    from oracle import learner as L

    torch.manual_seed(123)
    T, N = 6, 80
    buf, last_obs, last_priv = L.synthetic_rollout(T, N, seed=9, done_rate=0.05, timeout_rate=0.06)
    model = M.ActorCritic(12, 47, 14)
    with torch.no_grad():
        model.actor[6].weight.mul_(6.0)
        model.logstd.add_(torch.linspace(-0.3, 0.3, 12).view(1, 12))
        buf["actions"] = model.act(buf["obses"]).sample()
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    lr = 1e-3
    save = {"T": T, "N": N, "last_obs": last_obs.numpy(), "last_priv": last_priv.numpy()}
    save.update({"buf_" + k: v.numpy() for k, v in buf.items()})
    save.update({"sd0_" + k: v for k, v in sd0.items()})
    # --- the body of utils/runner.py:123-180, driven through the reference's own classes and helpers
    with torch.no_grad():
        old_dist = model.act(buf["obses"])
        old_logp = old_dist.log_prob(buf["actions"]).sum(dim=-1)
    save["old_logp"] = old_logp.numpy()
    save["old_mu"] = old_dist.loc.numpy()
    rewards = buf["rewards"].clone()
    for ep in range(2):
        values = model.est_value(buf["obses"], buf["privileged_obses"])
        last_values = model.est_value(last_obs, last_priv)
        with torch.no_grad():
            rewards[buf["time_outs"]] = values[buf["time_outs"]]
            adv = U.discount_values(rewards, buf["dones"] | buf["time_outs"], values, last_values, 0.995, 0.95)
            returns = values + adv
            advn = (adv - adv.mean()) / (adv.std() + 1e-8)
        value_loss = F.mse_loss(values, returns)
        dist = model.act(buf["obses"])
        logp = dist.log_prob(buf["actions"]).sum(dim=-1)
        actor_loss = U.surrogate_loss(old_logp, logp, advn)
        bound_loss = torch.clip(dist.loc - 1.0, min=0.0).square().mean() + torch.clip(dist.loc + 1.0, max=0.0).square().mean()
        entropy = dist.entropy().sum(dim=-1)
        loss = value_loss + actor_loss + 1.0 * bound_loss + (-0.01) * entropy.mean()
        opt.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone().numpy() for k, p in model.named_parameters()}
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        with torch.no_grad():
            kl = torch.sum(torch.log(dist.scale / old_dist.scale)
                           + 0.5 * (torch.square(old_dist.scale) + torch.square(dist.loc - old_dist.loc)) / torch.square(dist.scale) - 0.5, axis=-1)
            kl_mean = torch.mean(kl)
            if kl_mean > 0.01 * 2:
                lr = max(1e-5, lr / 1.5)
            elif kl_mean < 0.01 / 2:
                lr = min(1e-2, lr * 1.5)
            for pg in opt.param_groups:
                pg["lr"] = lr
        p = f"ep{ep}_"
        save.update({p + "values": values.detach().numpy(), p + "last_values": last_values.detach().numpy(), p + "adv": adv.numpy(),
                     p + "returns": returns.numpy(), p + "advn": advn.numpy(), p + "mu": dist.loc.detach().numpy(),
                     p + "value_loss": value_loss.item(), p + "actor_loss": actor_loss.item(), p + "bound_loss": bound_loss.item(),
                     p + "entropy": entropy.mean().item(), p + "kl": kl_mean.item(), p + "lr": lr, p + "rewards": rewards.numpy().copy()})
        save.update({p + "grad_" + k: v for k, v in grads.items()})
        save.update({p + "param_" + k: v.detach().clone().numpy() for k, v in model.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "learner_small.npz"), **save)
    print("learner: value_loss", save["ep0_value_loss"], "kl", save["ep1_kl"], "lr", save["ep1_lr"])
    # GAE alone, with edge cases (all done, none done, T = 1)
    g = torch.Generator().manual_seed(4)
    cases = {}
    for name, (T_, N_, pd) in {"a": (24, 64, 0.1), "b": (1, 5, 0.5), "c": (7, 3, 1.0), "d": (7, 3, 0.0)}.items():
        r = torch.rand(T_, N_, generator=g); v = torch.randn(T_, N_, generator=g); lv = torch.randn(N_, generator=g)
        d = torch.rand(T_, N_, generator=g) < pd
        a = U.discount_values(r, d, v, lv, 0.995, 0.95)
        cases.update({f"{name}_r": r.numpy(), f"{name}_v": v.numpy(), f"{name}_lv": lv.numpy(), f"{name}_d": d.numpy(), f"{name}_adv": a.numpy()})
    np.savez_compressed(os.path.join(OUT, "gae_small.npz"), **cases)


def gen_actor(ref):
    path = os.path.join(ref, "deploy", "models", "T1.pt")
    m = torch.jit.load(path, map_location="cpu")
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    names = sorted(sd.keys())
    print("T1.pt tensors:", names)
    g = torch.Generator().manual_seed(0)
    obs = torch.cat([torch.zeros(1, 47), torch.randn(63, 47, generator=g)])
    with torch.no_grad():
        mu = m(obs).numpy()
    out = {"obs": obs.numpy(), "mu": mu}
    for k, v in sd.items():  # TorchScript of model.actor: keys "0.weight" ... "6.bias"
        out["actor." + k] = v
    np.savez_compressed(os.path.join(OUT, "t1_actor_known_answer.npz"), **out)
    print("actor(zeros) =", mu[0])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = H.reference_dir()
    if ref is None:
        raise SystemExit("reference tree not found")
    mods = H.import_reference()
    if "--contacts-only" in sys.argv:
        gen_contacts(mods)
        raise SystemExit(0)
    gen_env(mods)
    gen_contacts(mods)
    gen_learner(mods)
    gen_actor(ref)
    print("fixtures written to", OUT)
