"""`envs` package of the drop-in surface: the reference's `from envs import *` (utils/runner.py:16) resolves T1 here."""
from booster_gym_b200.envs.t1 import T1  # noqa: F401
from booster_gym_b200.envs.base_task import BaseTask  # noqa: F401
