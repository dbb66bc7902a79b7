import isaacgym  # noqa: F401
from utils.runner import Runner

if __name__ == "__main__":
    runner = Runner(test=True)
    runner.play()
