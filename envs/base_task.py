from booster_gym_b200.envs.base_task import BaseTask  # noqa: F401
