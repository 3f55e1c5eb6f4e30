from nightmare_rl_b200.ppo import OnPolicyRunner  # noqa: F401
