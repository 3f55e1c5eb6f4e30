"""Shared helpers of the GPU parity tests."""
import numpy as np
import torch

from conftest import NMB
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import Batch
from nightmare_rl_b200.envcfg import build_envcfg
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from oracle import oracle as O

DEV = torch.device("cuda:0")
_cache = {}


def models():
    if "m" not in _cache:
        cm = mjcf.CompiledModel.load(NMB)
        _cache["m"] = (cm, _lib.Model(cm.to_bytes()), O.OracleModel(NMB))
    return _cache["m"]


def per_env_rel(a, b, floor=1e-3):
    """max |a-b| over the state vector, relative to the largest magnitude of that env's reference vector."""
    return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), floor)


def elem_rel(a, b, floor=1e-2):
    """element-wise |a-b| / max(|b|, floor), worst element of each env (per_env_rel can hide a large relative error
    on a small component behind the env's largest one)."""
    return (np.abs(a - b) / np.maximum(np.abs(b), floor)).max(axis=1)


def gpu_state(gb):
    return gb.qpos.cpu().numpy(), gb.qvel.cpu().numpy(), gb.warm.cpu().numpy()


def push_state(gb, q, v, w):
    gb.qpos.copy_(torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)))
    gb.qvel.copy_(torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)))
    gb.warm.copy_(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)))


def make_env_pair(n, seed, cfg=None, debug=False):
    cm, dm, om = models()
    cfg = cfg or NightmareV3Config()
    cfg.env.num_envs = n
    ec = build_envcfg(cfg, 0.008)
    ob = O.OracleBatch(om, n, seed=seed, envcfg=ec)
    gb = Batch(dm, n, DEV, seed=seed, envcfg=ec, debug=debug)
    return cfg, ob, gb


def sync_env_from_oracle(ob, gb):
    """Copy the oracle's complete env state (physics + carry buffers) into the GPU batch, rounded to fp32 on BOTH sides."""
    q, v, w = ob.get_state()
    q32, v32, w32 = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
    ob.set_state(q32, v32, w32)
    push_state(gb, q32, v32, w32)
    for name, dst in (("actions", gb.actions), ("dof_pos", gb.dof_pos), ("dof_vel", gb.dof_vel), ("commands", gb.commands),
                      ("episode_sums", gb.episode_sums)):
        val = ob.env_get(name).astype(np.float32)
        ob.env_set(name, val)
        dst.copy_(torch.from_numpy(val))
    gb.episode_length.copy_(torch.from_numpy(ob.env_get("ep_len").astype(np.int64)))
