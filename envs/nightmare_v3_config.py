from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO  # noqa: F401
