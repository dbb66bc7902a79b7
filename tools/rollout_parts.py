"""Live (CUDA-event) timing of the rollout components through the public API: policy, physics x10, post-physics, full step.
Note: timed in a Python loop, so each figure is max(GPU time, host issue cost) - see DESIGN.md (K4)."""
import copy, os, sys, numpy as np, torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ -> repo root
sys.path.insert(0, ROOT)
from booster_gym_b200.envs import T1
from booster_gym_b200.learner import Learner
from oracle import learner as L
cfg = yaml.load(open(os.path.join(ROOT, "envs", "T1.yaml")).read(), Loader=yaml.FullLoader)
n = int(os.environ.get("N", 4096)); cfg["env"]["num_envs"] = n; cfg["terrain"]["type"] = "plane"
env = T1(cfg); lrn = Learner(cfg, n, "cuda:0"); lrn.load_state_dict(L.init_params(0))
obs, infos = env.reset(); act = torch.zeros(n, 12, device="cuda")
def timeit(fn, k=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k * 1e3
print("policy_act us", timeit(lambda: lrn.act(obs, act)))
print("physics(10) us", timeit(lambda: env.physics(act, 10)))
print("post_physics us", timeit(lambda: env.post_physics()))
print("step us", timeit(lambda: env.step(act)))
