from booster_gym_b200.envs.t1 import T1  # noqa: F401
