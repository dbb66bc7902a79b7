from booster_gym_b200.utils.buffer import *  # noqa: F401,F403
from booster_gym_b200.utils.buffer import ExperienceBuffer  # noqa: F401,E402
