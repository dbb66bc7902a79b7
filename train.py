import isaacgym  # noqa: F401  (kept so the entry point reads like the reference's train.py)
from utils.runner import Runner

if __name__ == "__main__":
    runner = Runner(test=False)
    runner.train()
