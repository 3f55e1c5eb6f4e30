from nightmare_rl_b200.ppo import ActorCritic  # noqa: F401
