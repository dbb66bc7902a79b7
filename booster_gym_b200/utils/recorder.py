"""Recorder: TensorBoard / wandb scalars and checkpoints (mirror of utils/recorder.py:9-79).

Same constructor, directory layout (`logs/<timestamp>/{nn,summaries,config.yaml}`), scalar keys and `save()` format.
`record_episode_statistics` keeps the reference signature for callers that still feed per-step tensors, but the
Runner of this package uses `record_episode_summary`, fed once per iteration from the device-side accumulators of
the env (b200_t1_episode_stats) - the per-step `.item()` loops of the reference (utils/recorder.py:41-53) are gone.
"""
import os
import time

import torch
import yaml


class Recorder:

    def __init__(self, cfg):
        self.cfg = cfg
        name = time.strftime("%Y-%m-%d-%H-%M-%S", time.localtime())
        self.dir = os.path.join("logs", name)
        os.makedirs(self.dir, exist_ok=True)
        self.model_dir = os.path.join(self.dir, "nn")
        os.makedirs(self.model_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter

            self.writer = SummaryWriter(os.path.join(self.dir, "summaries"))
        except Exception as e:  # tensorboard not installed: keep training, say so once
            print(f"Recorder: TensorBoard unavailable ({e}); scalars are not written")
            self.writer = None
        self._wandb = None
        if self.cfg["runner"]["use_wandb"]:
            import wandb

            wandb.init(project=self.cfg["basic"]["task"], dir=self.dir, name=name, notes=self.cfg["basic"]["description"],
                       config=self.cfg)
            self._wandb = wandb

        self.episode_statistics = {}
        self.last_episode = {"steps": []}
        self.episode_steps = None

        with open(os.path.join(self.dir, "config.yaml"), "w") as file:
            yaml.dump(self.cfg, file)

    def _scalar(self, path, value, it):
        if self.writer is not None:
            self.writer.add_scalar(path, value, it)
        if self._wandb is not None:
            self._wandb.log({path: value}, step=it)

    @staticmethod
    def _path(key):
        return ("" if key in ("steps", "reward") else "episode/") + key

    def record_episode_summary(self, means, count, it):
        """means: {"reward", "steps", <term>...} averaged over the `count` episodes that ended this iteration"""
        for key, value in means.items():
            self._scalar(self._path(key), float(value) if count > 0 else 0.0, it)

    def record_episode_statistics(self, done, ep_info, it, write_record=False):
        """reference-compatible per-step path (host syncs; not used by this package's Runner)"""
        if self.episode_steps is None:
            self.episode_steps = torch.zeros_like(done, dtype=int)
        else:
            self.episode_steps += 1
        self.last_episode["steps"].extend(self.episode_steps[done].tolist())
        self.episode_steps[done] = 0
        for key, value in ep_info.items():
            if self.episode_statistics.get(key) is None:
                self.episode_statistics[key] = torch.zeros_like(value)
            self.episode_statistics[key] += value
            self.last_episode.setdefault(key, []).extend(self.episode_statistics[key][done].tolist())
            self.episode_statistics[key][done] = 0
        if write_record:
            for key, vals in self.last_episode.items():
                self._scalar(self._path(key), self._mean(vals), it)
                vals.clear()

    def record_statistics(self, statistics, it):
        for key, value in statistics.items():
            self._scalar(key, float(value), it)

    def save(self, model_dict, it):
        path = os.path.join(self.model_dir, "model_{}.pth".format(it))
        print("Saving model to {}".format(path))
        torch.save(model_dict, path)

    def _mean(self, data):
        return sum(data) / len(data) if len(data) else 0.0
