from nightmare_rl_b200.ppo import RolloutStorage  # noqa: F401
