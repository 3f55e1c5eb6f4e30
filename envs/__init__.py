"""Import shim with the reference's package name: ``from envs.nightmare_v3_env import NightmareV3Env`` etc. resolve to
the B200 implementation in ``nightmare_rl_b200.envs``."""
