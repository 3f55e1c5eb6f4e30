from nightmare_rl_b200.ppo import PPO  # noqa: F401
