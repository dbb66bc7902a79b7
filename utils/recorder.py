from booster_gym_b200.utils.recorder import *  # noqa: F401,F403
from booster_gym_b200.utils.recorder import Recorder  # noqa: F401,E402
