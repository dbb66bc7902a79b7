"""Runner: the PPO training / play loop behind the interface of utils/runner.py:19-245.

Same constructor (`Runner(test)`; argparse flags `--task --checkpoint --num_envs --headless --sim_device --rl_device
--seed --max_iterations`, YAML `envs/<task>.yaml`), same `train()` / `play()`, same checkpoint dict, same logged scalar
keys.  What differs is WHERE the arithmetic runs: every tensor operation between `train()` entry and the optimizer step
is a kernel of libb200t1.so launched on torch's current stream (policy forward + sampling, env step, GAE, the epoch's
forward / losses / backward, clip + Adam + KL learning-rate rule); the host reads device scalars once per iteration.

Multi-GPU: launched with torchrun, one process per GPU; rank r simulates envs [r*N, (r+1)*N) of a global
layout of world*N envs; per epoch the advantage moments (3 doubles) are sum-all-reduced before the loss and the flat
gradient buffer plus the loss sums (for the KL rule) after the backward pass; every rank then applies the identical
clip + Adam step.  NCCL over NVLink; nothing else crosses ranks.
"""
import argparse
import glob
import os
import random
import time

import numpy as np
import torch
import yaml

from .. import _abi
from ..envs import T1  # noqa: F401  (task classes are resolved by name)
from ..learner import Learner
from .buffer import ExperienceBuffer
from .model import ActorCritic
from .recorder import Recorder

TASKS = {"T1": T1}


class FlatAdam:
    """torch.optim.Adam-shaped facade over the learner's flat Adam buffers: `state_dict()` / `load_state_dict()` use the
    layout torch.optim.Adam(model.parameters()) produces for the reference model, so `.pth` files interchange."""

    def __init__(self, model, learner):
        self.model, self.learner = model, learner

    @property
    def param_groups(self):
        lr = float(self.learner.scalars[_abi.SC["LR"]].item())
        n = len(list(self.model.parameters()))
        return [dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                     capturable=False, differentiable=False, fused=None, params=list(range(n)))]

    def state_dict(self):
        m, v = self.learner.views(self.learner.adam_m), self.learner.views(self.learner.adam_v)
        step = self.learner.scalars[_abi.SC["ADAM_STEP"]].detach().cpu().clone()
        state = {}
        if step.item() > 0:
            for i, (name, _) in enumerate(self.model.named_parameters()):
                state[i] = {"step": step.clone(), "exp_avg": m[name].clone().reshape(getattr_path(self.model, name).shape),
                            "exp_avg_sq": v[name].clone().reshape(getattr_path(self.model, name).shape)}
        return {"state": state, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        m, v = self.learner.views(self.learner.adam_m), self.learner.views(self.learner.adam_v)
        names = [n for n, _ in self.model.named_parameters()]
        for i, st in sd["state"].items():
            name = names[int(i)]
            m[name].copy_(st["exp_avg"].to(m[name].device).reshape(m[name].shape))
            v[name].copy_(st["exp_avg_sq"].to(v[name].device).reshape(v[name].shape))
            self.learner.scalars[_abi.SC["ADAM_STEP"]] = float(st["step"])
        if sd.get("param_groups"):
            self.learner.set_lr(sd["param_groups"][0]["lr"])


def getattr_path(obj, path):
    for part in path.split("."):
        obj = getattr(obj, part)
    return obj


class Runner:

    def __init__(self, test=False, argv=None, cfg_overrides=None):
        self.test = test
        self._get_args(argv)
        self._update_cfg_from_args()
        for section, values in (cfg_overrides or {}).items():  # programmatic use (bench.py, tests): {"terrain": {"type": "plane"}}
            self.cfg[section].update(values)
        self._init_distributed()
        self._set_seed()
        task_class = TASKS.get(self.cfg["basic"]["task"])
        if task_class is None:
            raise NameError(f"name '{self.cfg['basic']['task']}' is not defined")
        n = self.cfg["env"]["num_envs"]
        self.env = task_class(self.cfg, env_index_base=self.rank * n, total_envs=self.world_size * n)

        self.device = self.cfg["basic"]["rl_device"]
        if torch.device(self.device) != torch.device(self.env.device):
            raise ValueError("rl_device must equal sim_device: rollout tensors never leave the GPU in this implementation")
        self.learning_rate = self.cfg["algorithm"]["learning_rate"]
        self.learner = Learner(self.cfg, self.env.num_envs, self.device, world_size=self.world_size,
                               env_base=self.rank * n, learning_rate=self.learning_rate, seed=self.cfg["basic"]["seed"])
        self.model = ActorCritic(self.env.num_actions, self.env.num_obs, self.env.num_privileged_obs).bind(self.learner)
        if self.world_size > 1:
            torch.distributed.broadcast(self.learner.params, src=0)
            if os.environ.get("B200_PEER_EXCHANGE", "1") != "0":
                self.learner.bind_peers()   # NVLink peer-memory exchange inside the learner kernels (raises if it cannot be bound)
        self.optimizer = FlatAdam(self.model, self.learner)
        self._load()

        self.buffer = ExperienceBuffer(self.cfg["runner"]["horizon_length"], self.env.num_envs, self.device)
        self.buffer.add_buffer("actions", (self.env.num_actions,))
        self.buffer.add_buffer("obses", (self.env.num_obs,))
        self.buffer.add_buffer("privileged_obses", (self.env.num_privileged_obs,))
        self.buffer.add_buffer("rewards", ())
        self.buffer.add_buffer("dones", (), dtype=bool)
        self.buffer.add_buffer("time_outs", (), dtype=bool)

    # ---- configuration (same flags / YAML handling as the reference) ---------------------------------------------
    def _get_args(self, argv=None):
        parser = argparse.ArgumentParser()
        parser.add_argument("--task", required=True, type=str, help="Name of the task to run.")
        parser.add_argument("--checkpoint", type=str, help="Path of the model checkpoint to load. Overrides config file if provided.")
        parser.add_argument("--num_envs", type=int, help="Number of environments to create. Overrides config file if provided.")
        parser.add_argument("--headless", type=bool, help="Run headless without creating a viewer window. Overrides config file if provided.")
        parser.add_argument("--sim_device", type=str, help="Device for physics simulation. Overrides config file if provided.")
        parser.add_argument("--rl_device", type=str, help="Device for the RL algorithm. Overrides config file if provided.")
        parser.add_argument("--seed", type=int, help="Random seed. Overrides config file if provided.")
        parser.add_argument("--max_iterations", type=int, help="Maximum number of training iterations. Overrides config file if provided.")
        self.args = parser.parse_args(argv)

    def _update_cfg_from_args(self):
        cfg_file = os.path.join("envs", "{}.yaml".format(self.args.task))
        with open(cfg_file, "r", encoding="utf-8") as f:
            self.cfg = yaml.load(f.read(), Loader=yaml.FullLoader)
        for arg, val in vars(self.args).items():
            if val is None:
                continue
            section = "env" if arg == "num_envs" else "basic"
            self.cfg[section][arg] = val
        if not self.test:
            self.cfg["viewer"]["record_video"] = False

    def _init_distributed(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world_size > 1:
            local = int(os.environ.get("LOCAL_RANK", str(self.rank)))
            dev = f"cuda:{local}"
            self.cfg["basic"]["sim_device"] = dev
            self.cfg["basic"]["rl_device"] = dev
            torch.cuda.set_device(local)
            if not torch.distributed.is_initialized():
                torch.distributed.init_process_group(backend="nccl", device_id=torch.device(dev))

    def _set_seed(self):
        if self.cfg["basic"]["seed"] == -1:
            self.cfg["basic"]["seed"] = np.random.randint(0, 10000)
        seed = self.cfg["basic"]["seed"]
        if self.rank == 0:
            print("Setting seed: {}".format(seed))
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        os.environ["PYTHONHASHSEED"] = str(seed)
        torch.cuda.manual_seed_all(seed)

    def _load(self):
        ckpt = self.cfg["basic"]["checkpoint"]
        if not ckpt:
            return
        if ckpt == "-1" or ckpt == -1:
            ckpt = sorted(glob.glob(os.path.join("logs", "**/*.pth"), recursive=True), key=os.path.getmtime)[-1]
            self.cfg["basic"]["checkpoint"] = ckpt
        print("Loading model from {}".format(ckpt))
        model_dict = torch.load(ckpt, map_location=self.device, weights_only=True)
        self.model.load_state_dict(model_dict["model"], strict=False)
        try:
            self.env.curriculum_prob = model_dict["curriculum"]
        except Exception as e:
            print(f"Failed to load curriculum: {e}")
        try:
            self.optimizer.load_state_dict(model_dict["optimizer"])
        except Exception as e:
            print(f"Failed to load optimizer: {e}")

    # ---- the loop ----------------------------------------------------------------------------------------------------
    def rollout(self, obs, privileged_obs):
        """utils/runner.py:106-121: horizon_length x [store obs, act, env.step, store transition].  The transition is stored by the
        producing kernels themselves: the policy samples into the `actions` row, the env step writes reward / done into their rows
        and the NEXT observation into row n + 1 (`T1.step(_out=...)`); only row 0 of the observations and the time-outs (whose tensor
        is rebound only on steps with a reset, SURVEY 8a note 1) are copied.  The last step writes the env's own buffers, which
        are what this returns - as the reference's loop does."""
        buf, env, lrn = self.buffer, self.env, self.learner
        horizon = self.cfg["runner"]["horizon_length"]
        buf.update_data("obses", 0, obs)
        buf.update_data("privileged_obses", 0, privileged_obs)
        for n in range(horizon):
            act = buf["actions"][n]
            # mu = actor(obs); act = mu + sigma * eps, written into the buffer row (the parameters only change in update(): every
            # step but the first may reuse the weight operands of the tensor-core act path)
            lrn.act(obs, act, reuse_weights=n > 0)
            last = n == horizon - 1
            out = (env.obs_buf if last else buf["obses"][n + 1], env.privileged_obs_buf if last else buf["privileged_obses"][n + 1],
                   buf["rewards"][n], buf.row("dones", n))
            obs, rew, done, infos = env.step(act, _device_counter=True, _out=out)
            privileged_obs = infos["privileged_obs"]
            buf.update_data("time_outs", n, infos["time_outs"])
        return obs, privileged_obs

    def rollout_graphed(self, obs, privileged_obs):
        """the same rollout replayed from a CUDA graph: 24 x (6 buffer writes + policy + physics + post-physics) are ~220 launches whose
        host-side issue cost (Python + ctypes + torch dispatch) exceeds their GPU time at 4096 envs; RNG / step counters live on the
        device, so a replay draws fresh noise.  The first call runs eagerly (a normal rollout) and then captures."""
        if getattr(self, "_rollout_graph", None) is None:
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                out = self.rollout(obs, privileged_obs)          # eager: this IS iteration 0's rollout
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._rollout_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._rollout_graph):          # recorded, not executed
                self._rollout_out = self.rollout(out[0], out[1])
            return out
        self._rollout_graph.replay()
        return self._rollout_out

    def update(self, obs, privileged_obs):
        """utils/runner.py:123-185: old distribution, then mini_epochs x [values, GAE, losses, backward, clip, Adam, KL]"""
        buf, lrn = self.buffer, self.learner
        lrn.old_dist(buf["obses"], buf["privileged_obses"], buf["actions"])
        dones, touts = buf.raw("dones"), buf.raw("time_outs")
        multi = self.world_size > 1 and not lrn.peers_bound   # peers_bound: the library exchanges over NVLink peer memory itself
        for _ in range(self.cfg["runner"]["mini_epochs"]):
            lrn.epoch_a(buf["rewards"], dones, touts, obs, privileged_obs)
            if multi:
                torch.distributed.all_reduce(lrn.dstats[0:4])
            lrn.epoch_b(buf["actions"])
            if multi:
                torch.distributed.all_reduce(lrn.grads)
                torch.distributed.all_reduce(lrn.dstats[4:10])
            lrn.apply()

    def update_graphed(self, obs, privileged_obs):
        """the same update replayed from a CUDA graph: 20 epochs x ~17 launches with ~1-3 us of dependent-launch latency each.  lr,
        Adam step, the KL rule and (multi-GPU, peer-memory exchange) the exchange sequence numbers live on the device, every pointer
        is a fixed buffer, so a replay IS the next update.  Only the NCCL protocol (B200_PEER_EXCHANGE=0), which interleaves
        collectives issued from Python, stays eager.  The first call runs eagerly (warm-up), the second captures."""
        if (self.world_size > 1 and not self.learner.peers_bound) or os.environ.get("B200_UPDATE_GRAPH", "1") == "0":
            return self.update(obs, privileged_obs)
        key = (obs.data_ptr(), privileged_obs.data_ptr())
        g = getattr(self, "_update_graph", None)
        if g is not None and self._update_graph_key == key:
            g.replay()
            return
        if getattr(self, "_update_warm", None) != key:
            self._update_warm = key                 # first sight of these buffers: a plain update (also the capture warm-up)
            return self.update(obs, privileged_obs)
        torch.cuda.synchronize(self.device)
        self._update_graph = torch.cuda.CUDAGraph()
        self._update_graph_key = key
        with torch.cuda.graph(self._update_graph):  # recorded, not executed
            self.update(obs, privileged_obs)
        self._update_graph.replay()

    def train(self, on_iteration=None):
        """utils/runner.py:99-215.  `on_iteration(it, episode_means, episode_count, scalars)` (not in the reference surface) is
        called once per iteration with what the Recorder is fed - tools/learning_curve.py and the learning test use it.

        Nothing on the device waits for the host (lr, KL rule, Adam step, RNG / step counters are device-resident), so the loop is
        pipelined: the scalars / episode statistics / curriculum levels of iteration i are copied to pinned host memory in stream
        order, the graphs of iteration i + 1 are launched, and only then the host waits for iteration i's copies and logs them
        (TensorBoard, callbacks) while the GPU is busy.  A checkpoint needs the parameters OF its iteration and drains the pipeline."""
        self.recorder = Recorder(self.cfg) if self.rank == 0 else None
        obs, infos = self.env.reset()
        privileged_obs = infos["privileged_obs"]
        SC = _abi.SC
        cur_multi = self.world_size > 1 and bool(self.cfg["commands"].get("curriculum"))
        use_graph = os.environ.get("B200_ROLLOUT_GRAPH", "1") != "0"
        pipelined = os.environ.get("B200_TRAIN_PIPELINE", "1") != "0"
        horizon = self.cfg["runner"]["horizon_length"]
        n_stats = len(self.env.reward_names) + 3
        slots = [{"scalars": torch.empty(SC["COUNT"], dtype=torch.float32).pin_memory(), "stats": torch.zeros(n_stats, dtype=torch.float64).pin_memory(),
                  "levels": torch.zeros(4, dtype=torch.float32).pin_memory(), "event": torch.cuda.Event(), "it": -1} for _ in range(2)]
        max_it = self.cfg["basic"]["max_iterations"]

        def publish(slot):
            """host side of one finished iteration: wait for its copies, then callbacks / Recorder"""
            slot["event"].synchronize()
            it, sc = slot["it"], slot["scalars"]
            epochs = max(1.0, sc[SC["EPOCHS"]].item())
            self.learning_rate = sc[SC["LR"]].item()
            ep_means, ep_count = self.env.decode_episode_stats(slot["stats"])
            if on_iteration is not None:
                on_iteration(it, ep_means, ep_count, sc)
            if self.recorder is not None:
                lv = slot["levels"]
                self.recorder.record_episode_summary(ep_means, ep_count, it)
                self.recorder.record_statistics(
                    {
                        "value_loss": sc[SC["SUM_VALUE_LOSS"]].item() / epochs,
                        "actor_loss": sc[SC["SUM_ACTOR_LOSS"]].item() / epochs,
                        "bound_loss": sc[SC["SUM_BOUND_LOSS"]].item() / epochs,
                        "entropy": sc[SC["SUM_ENTROPY"]].item() / epochs,
                        "kl_mean": sc[SC["KL"]].item(),
                        "lr": self.learning_rate,
                        "curriculum/mean_lin_vel_level": lv[0].item(),
                        "curriculum/mean_ang_vel_level": lv[1].item(),
                        "curriculum/max_lin_vel_level": lv[2].item(),
                        "curriculum/max_ang_vel_level": lv[3].item(),
                    },
                    it,
                )
                print("epoch: {}/{}".format(it + 1, max_it))

        pending = None
        step_base = self.env.common_step_counter
        for it in range(max_it):
            if cur_multi:
                cur_before = self.env.curriculum_prob.clone()
            obs, privileged_obs = (self.rollout_graphed if use_graph else self.rollout)(obs, privileged_obs)
            if cur_multi:
                # every rank raised its own copy of the command-curriculum grid from its own envs' successes; merge once per
                # iteration: sum of the ranks' increments on top of the common starting grid, clamped like envs/t1.py:413
                # (SURVEY 8e "other cross-rank state"; a single-process run clamps after every step instead - same fixed point)
                delta = self.env.curriculum_prob - cur_before
                torch.distributed.all_reduce(delta)
                self.env.curriculum_prob = torch.clamp(cur_before + delta, max=1.0)
            self.learner.scalars[SC["SUM_VALUE_LOSS"]:SC["EPOCHS"] + 1].zero_()
            (self.update_graphed if use_graph else self.update)(obs, privileged_obs)
            # ---- this iteration's results -> pinned host memory, in stream order (the next iteration overwrites the device copies)
            slot = slots[it & 1]
            slot["it"] = it
            slot["scalars"].copy_(self.learner.scalars, non_blocking=True)
            self.env.episode_stats_async(slot["stats"])
            lvl = torch.abs(self.env.env_curriculum_level).float()
            slot["levels"].copy_(torch.cat([lvl.mean(0), lvl.max(0).values]), non_blocking=True)
            slot["event"].record(torch.cuda.current_stream(self.device))
            self.env.common_step_counter = step_base + (it + 1) * horizon   # (the device counter advances once per env step: no read-back)
            # ---- the previous iteration is logged while this one runs
            if pending is not None:
                publish(pending)
            pending = slot
            save_now = self.recorder is not None and (it + 1) % self.cfg["runner"]["save_interval"] == 0
            if not pipelined or save_now or it == max_it - 1:
                publish(pending)
                pending = None
            if save_now:
                self.recorder.save(
                    {"model": self.model.state_dict(), "optimizer": self.optimizer.state_dict(),
                     "curriculum": self.env.curriculum_prob},
                    it + 1,
                )

    def play(self, max_steps=None):
        obs, infos = self.env.reset()
        act = torch.empty(self.env.num_envs, self.env.num_actions, dtype=torch.float32, device=self.device)
        steps = 0
        if self.cfg["viewer"]["record_video"]:
            print("record_video is ignored: this implementation has no renderer")
        while max_steps is None or steps < max_steps:
            self.learner.act(obs, act, deterministic=True)   # dist.loc (utils/runner.py:226-227)
            obs, rew, done, infos = self.env.step(act)
            steps += 1
