"""Flatten a ``NightmareV3Config`` into the scalar table the step kernels read.

Mirrors what ``NightmareV3Env.__init__`` derives from the config (reference
``envs/nightmare_v3_env.py:80-137``): ``dt``, ``max_episode_length``, the command
resampling period, reward scales with zeros dropped and the rest multiplied by ``dt``, the reward
evaluation order (alphabetical, ``termination`` last) and the 66-entry observation-noise vector."""
from __future__ import annotations

import ctypes
import math

import numpy as np

from .envs.helpers import class_to_dict

NUM_DOF = 18
NUM_OBS = 66
# alphabetical ids shared by the CUDA kernels, the C-ABI header and the oracle
REWARD_TERMS = ("action_rate", "ang_vel_xy", "base_height", "body_contact_forces", "collision", "default_position",
                "dof_acc", "dof_vel", "feet_air_time", "feet_contact_forces", "feet_stumble", "lin_vel_z", "orientation",
                "stand_still", "termination", "torques", "tracking_ang_vel", "tracking_lin_vel")
NREW = len(REWARD_TERMS)
_NO_FUNCTION = ("collision", "feet_stumble")   # scales exist, `_reward_*` functions do not (env.py:137 raises)


class EnvCfgStruct(ctypes.Structure):
    """Binary layout of ``nm_envcfg`` (include/nightmare_b200.h) and ``nmo_envcfg`` (oracle)."""
    _fields_ = [
        ("decimation", ctypes.c_int32), ("num_actions", ctypes.c_int32),
        ("tibia_contact_mode", ctypes.c_int32), ("body_contact_mode", ctypes.c_int32),
        ("add_noise", ctypes.c_int32), ("resample_period", ctypes.c_int32),
        ("strict_reference", ctypes.c_int32), ("pad0", ctypes.c_int32),
        ("action_scale", ctypes.c_double), ("clip_actions", ctypes.c_double), ("p_gain", ctypes.c_double), ("clip_obs", ctypes.c_double),
        ("default_pos", ctypes.c_double * NUM_DOF),
        ("obs_lin_vel", ctypes.c_double), ("obs_ang_vel", ctypes.c_double), ("obs_dof_pos", ctypes.c_double), ("obs_dof_vel", ctypes.c_double),
        ("max_lin_vel_x", ctypes.c_double), ("max_ang_vel", ctypes.c_double),
        ("max_episode_length", ctypes.c_double), ("max_episode_length_s", ctypes.c_double),
        ("termination_contact_force", ctypes.c_double), ("tibia_max_contact_force", ctypes.c_double), ("body_max_contact_force", ctypes.c_double),
        ("tracking_sigma", ctypes.c_double), ("base_height_target", ctypes.c_double), ("max_contact_force", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("rew_scale", ctypes.c_double * NREW),
        ("noise_vec", ctypes.c_double * NUM_OBS),
    ]


def reward_table(cfg, dt: float):
    """(scale*dt per term id, active names in accumulation order, episode-sum key order)."""
    scales = class_to_dict(cfg.rewards.scales)
    table = np.zeros(NREW)
    active = {}
    for name, s in scales.items():
        if s == 0:
            continue
        if name not in REWARD_TERMS or name in _NO_FUNCTION:
            # same failure the reference has at construction (getattr on a missing _reward_* method)
            raise AttributeError(f"'NightmareV3Env' object has no attribute '_reward_{name}'")
        active[name] = s * dt
        table[REWARD_TERMS.index(name)] = s * dt
    order = [n for n in active if n != "termination"]
    return table, order, list(active.keys())


def noise_vector(cfg) -> np.ndarray:
    """The literal 66-vector of env.py:109-119 (its 12-DoF slice boundaries are kept, quirk Q8)."""
    v = np.zeros(NUM_OBS)
    ns, lvl, sc = cfg.noise.noise_scales, cfg.noise.noise_level, cfg.normalization.obs_scales
    v[:3] = ns.lin_vel * lvl * sc.lin_vel
    v[3:6] = ns.ang_vel * lvl * sc.ang_vel
    v[6:9] = ns.gravity * lvl
    v[12:24] = ns.dof_pos * lvl * sc.dof_pos
    v[24:36] = ns.dof_vel * lvl * sc.dof_vel
    return v


def build_envcfg(cfg, timestep: float) -> EnvCfgStruct:
    dt = timestep * cfg.control.decimation
    s = EnvCfgStruct()
    s.decimation = int(cfg.control.decimation)
    s.num_actions = int(cfg.env.num_actions)
    s.tibia_contact_mode = int(cfg.env.tibia_contact_mode)
    s.body_contact_mode = int(cfg.env.body_contact_mode)
    s.add_noise = int(bool(cfg.noise.add_noise))
    # int32 field of the C ABI: a resampling time beyond ~1e7 s (play.py uses 1e9 to switch resampling off) saturates instead of wrapping
    s.resample_period = int(min(cfg.commands.resampling_time / dt, 2 ** 31 - 1))
    # quirk Q10 switch: 1 (default) latches extras only on steps where an env reset, like the reference; 0 refreshes
    # extras['time_outs'] every step (cfg.env.strict_reference = False)
    s.strict_reference = 1 if getattr(cfg.env, "strict_reference", True) else 0
    s.action_scale = cfg.control.action_scale
    s.clip_actions = cfg.normalization.clip_actions
    s.p_gain = cfg.control.p_gain
    s.clip_obs = cfg.normalization.clip_observations
    dp = np.asarray(cfg.control.default_pos, dtype=np.float64)
    if dp.shape != (NUM_DOF,):
        raise ValueError("control.default_pos must have 18 entries")
    s.default_pos[:] = dp.tolist()
    o = cfg.normalization.obs_scales
    s.obs_lin_vel, s.obs_ang_vel, s.obs_dof_pos, s.obs_dof_vel = o.lin_vel, o.ang_vel, o.dof_pos, o.dof_vel
    s.max_lin_vel_x = cfg.commands.ranges.max_lin_vel_x
    s.max_ang_vel = cfg.commands.ranges.max_ang_vel
    s.max_episode_length_s = cfg.env.episode_length_s
    s.max_episode_length = math.ceil(cfg.env.episode_length_s / dt)
    s.termination_contact_force = cfg.env.termination_contact_force
    s.tibia_max_contact_force = cfg.env.tibia_max_contact_force
    s.body_max_contact_force = cfg.env.body_max_contact_force
    s.tracking_sigma = cfg.rewards.tracking_sigma
    s.base_height_target = cfg.rewards.base_height_target
    s.max_contact_force = cfg.rewards.max_contact_force
    s.dt = dt
    table, _, _ = reward_table(cfg, dt)
    s.rew_scale[:] = table.tolist()
    s.noise_vec[:] = noise_vector(cfg).tolist()
    return s
