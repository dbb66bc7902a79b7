from booster_gym_b200.utils.runner import *  # noqa: F401,F403
from booster_gym_b200.utils.runner import Runner, FlatAdam  # noqa: F401,E402
