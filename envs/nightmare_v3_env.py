from nightmare_rl_b200.envs.nightmare_v3_env import *  # noqa: F401,F403
from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env  # noqa: F401
