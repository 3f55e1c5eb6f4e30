from nightmare_rl_b200.envs.base_config import BaseConfig  # noqa: F401
