"""The reference's entry point (train.py -> utils/runner.py:99-215) through the drop-in Runner on the GPU: a few complete iterations
(CUDA-graph rollout + 20-epoch update + device episode statistics + checkpoint), and the graph-replayed rollout against the eager one."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _runner(tmp_path, **over):
    from booster_gym_b200.utils.runner import Runner

    cwd = os.getcwd()
    os.chdir(ROOT)          # like the reference, the Runner reads envs/<task>.yaml relative to the working directory ...
    try:
        ov = {"terrain": {"type": "plane"}, "runner": {"use_wandb": False, "save_interval": 2}, "basic": {"max_iterations": 3}}
        for k, v in over.items():
            ov.setdefault(k, {}).update(v)
        runner = Runner(test=False, argv=["--task", "T1", "--num_envs", "512", "--headless", "True"], cfg_overrides=ov)
    finally:
        os.chdir(cwd)
    os.chdir(tmp_path)      # ... and train() writes logs/ relative to it
    return runner, cwd


def test_train_runs_three_iterations_and_writes_a_checkpoint(tmp_path):
    from booster_gym_b200 import _abi

    runner, cwd = _runner(tmp_path)
    try:
        p0 = runner.learner.params.clone()
        runner.train()
        torch.cuda.synchronize()
        assert torch.isfinite(runner.learner.params).all() and not torch.equal(p0, runner.learner.params)
        assert runner.learner.scalars[_abi.SC["ADAM_STEP"]].item() == 3 * runner.cfg["runner"]["mini_epochs"]
        assert getattr(runner, "_rollout_graph", None) is not None           # iterations 1.. were graph replays
        assert runner.env.common_step_counter == runner.env.counters()[1] == 3 * runner.cfg["runner"]["horizon_length"]   # host mirror of the device counter
        ckpts = [os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".pth")]
        assert len(ckpts) == 1
        ck = torch.load(ckpts[0], map_location="cpu", weights_only=True)
        assert set(ck) == {"model", "optimizer", "curriculum"} and ck["curriculum"].shape == (21, 21)
    finally:
        os.chdir(cwd)


def test_graph_replay_equals_eager_rollout(tmp_path):
    """same seed, same parameters: 2 rollouts eagerly vs 1 eager + 1 graph replay produce identical buffers (device-side RNG / step counters)"""
    outs = []
    for graphed in (False, True):
        runner, cwd = _runner(tmp_path)
        try:
            obs, infos = runner.env.reset()
            priv = infos["privileged_obs"]
            roll = runner.rollout_graphed if graphed else runner.rollout
            obs, priv = roll(obs, priv)
            obs, priv = roll(obs, priv)
            torch.cuda.synchronize()
            outs.append({k: runner.buffer[k].clone() for k in ("actions", "obses", "rewards", "dones", "time_outs")} | {"last_obs": obs.clone()})
        finally:
            os.chdir(cwd)
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_rollout_storage_matches_the_reference_loop(tmp_path):
    """Runner.rollout lets the kernels write the transition straight into the rollout storage (T1.step(_out=...)); the buffers must
    equal those of the reference's own loop shape (utils/runner.py:106-121: copy obs, act, step into the env's buffers, copy the rest)."""
    outs = []
    for direct in (True, False):
        runner, cwd = _runner(tmp_path)
        try:
            env, buf, lrn = runner.env, runner.buffer, runner.learner
            obs, infos = env.reset()
            priv = infos["privileged_obs"]
            for _ in range(2):
                if direct:
                    obs, priv = runner.rollout(obs, priv)
                else:
                    for n in range(runner.cfg["runner"]["horizon_length"]):
                        buf.update_data("obses", n, obs)
                        buf.update_data("privileged_obses", n, priv)
                        act = buf["actions"][n]
                        lrn.act(obs, act)
                        obs, rew, done, infos = env.step(act, _device_counter=True)
                        priv = infos["privileged_obs"]
                        buf.update_data("rewards", n, rew)
                        buf.update_data("dones", n, done)
                        buf.update_data("time_outs", n, infos["time_outs"])
            torch.cuda.synchronize()
            outs.append({k: buf[k].clone() for k in ("actions", "obses", "privileged_obses", "rewards", "dones", "time_outs")}
                        | {"last_obs": obs.clone(), "last_priv": priv.clone()})
        finally:
            os.chdir(cwd)
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
    assert outs[0]["dones"].any() and outs[0]["rewards"].abs().sum() > 0


def test_graph_replay_equals_eager_update(tmp_path):
    """Runner.update_graphed: the 20-epoch update replayed from a CUDA graph against the same update launched kernel by kernel, from the
    same parameters / Adam state / lr / rollout storage.  (The head and bias gradients are float atomics, so two runs agree to rounding,
    not to the bit: the difference is held to 1e-3 of the parameter change of the update.)"""
    runner, cwd = _runner(tmp_path)
    try:
        lrn = runner.learner
        obs, infos = runner.env.reset()
        priv = infos["privileged_obs"]
        for _ in range(2):                      # first sight of the buffers: eager; second: capture + replay
            obs, priv = runner.rollout_graphed(obs, priv)
            runner.update_graphed(obs, priv)
        assert getattr(runner, "_update_graph", None) is not None
        obs, priv = runner.rollout_graphed(obs, priv)
        torch.cuda.synchronize()
        snap = [t.clone() for t in (lrn.params, lrn.adam_m, lrn.adam_v, lrn.scalars, lrn.dstats, runner.buffer["rewards"])]

        def restore():
            for t, s in zip((lrn.params, lrn.adam_m, lrn.adam_v, lrn.scalars, lrn.dstats, runner.buffer["rewards"]), snap):
                t.copy_(s)

        runner.update_graphed(obs, priv)        # a replay
        torch.cuda.synchronize()
        a, sa = lrn.params.clone(), lrn.scalars.clone()
        restore()
        runner.update(obs, priv)                # eager
        torch.cuda.synchronize()
        b, sb = lrn.params.clone(), lrn.scalars.clone()
        moved = (b - snap[0]).abs().max().item()
        assert moved > 0 and torch.isfinite(a).all()
        assert (a - b).abs().max().item() <= 1e-3 * moved, ((a - b).abs().max().item(), moved)
        from booster_gym_b200 import _abi

        assert sa[_abi.SC["ADAM_STEP"]].item() == sb[_abi.SC["ADAM_STEP"]].item() == snap[3][_abi.SC["ADAM_STEP"]].item() + runner.cfg["runner"]["mini_epochs"]
        assert abs(sa[_abi.SC["LR"]].item() - sb[_abi.SC["LR"]].item()) <= 1e-6 * sb[_abi.SC["LR"]].item()
    finally:
        os.chdir(cwd)


def test_pipelined_logging_reports_every_iteration_in_order(tmp_path):
    """Runner.train logs iteration i while the graphs of iteration i + 1 run (pinned-memory snapshots + events): the callback must
    still see every iteration once, in order, with that iteration's scalars (Adam step = 20 (it + 1)), and the same numbers as the
    unpipelined loop (B200_TRAIN_PIPELINE=0) up to the float-atomic noise of the gradients."""
    from booster_gym_b200 import _abi

    seen = {}
    for mode in ("1", "0"):
        os.environ["B200_TRAIN_PIPELINE"] = mode
        try:
            runner, cwd = _runner(tmp_path, basic={"max_iterations": 5})
            rows = []
            try:
                runner.train(on_iteration=lambda it, means, count, sc: rows.append(
                    (it, int(sc[_abi.SC["ADAM_STEP"]].item()), float(sc[_abi.SC["LR"]].item()), count, means["steps"])))
            finally:
                os.chdir(cwd)
        finally:
            os.environ.pop("B200_TRAIN_PIPELINE", None)
        assert [r[0] for r in rows] == list(range(5))
        assert [r[1] for r in rows] == [runner.cfg["runner"]["mini_epochs"] * (i + 1) for i in range(5)]
        seen[mode] = rows
    for a, b in zip(seen["1"], seen["0"]):
        assert a[3] == b[3] and abs(a[4] - b[4]) <= 1e-9 * max(1.0, abs(b[4]))       # iteration 0's episodes are identical (same rollout)
        break
    assert sum(r[3] for r in seen["1"]) > 0
