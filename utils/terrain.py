from booster_gym_b200.utils.terrain import *  # noqa: F401,F403
from booster_gym_b200.utils.terrain import Terrain  # noqa: F401,E402
