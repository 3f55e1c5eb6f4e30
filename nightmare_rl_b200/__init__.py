"""nightmare_rl_b200 — B200-native batched environment step for the Nightmare v3 hexapod.

Only what the hot path needs lives here: the host model compiler (``mjcf``/``meshproc``), the
ctypes binding of the C-ABI CUDA library (``_lib``) and the ``NightmareV3Env`` drop-in (``envs``).
"""
__version__ = "0.1.0"
