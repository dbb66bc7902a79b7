from booster_gym_b200.utils.utils import *  # noqa: F401,F403
from booster_gym_b200.utils.utils import apply_randomization, discount_values, surrogate_loss  # noqa: F401,E402
