from booster_gym_b200.utils.model import *  # noqa: F401,F403
from booster_gym_b200.utils.model import ActorCritic  # noqa: F401,E402
