"""Golden fixtures (tests/golden/*.npz, written by tools/make_golden.py from the oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA env tracks them (see test_gpu_env.py)."""
import os

import numpy as np

from conftest import ROOT
from nightmare_rl_b200.envcfg import build_envcfg
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from oracle import oracle as O

GOLD = os.path.join(ROOT, "tests", "golden")


def _replay(oracle_model, g, n, seed, steps):
    cfg = NightmareV3Config()
    cfg.env.num_envs = n
    b = O.OracleBatch(oracle_model, n, seed=seed, envcfg=build_envcfg(cfg, 0.008))
    b.env_reset_idx(np.arange(n))
    if "ep0" in g:
        b.env_set("ep_len", g["ep0"])
    for t in range(steps):
        obs, rew, done, tout, _, _ = b.env_step(g["actions"][t].reshape(n, -1))
        assert np.array_equal(done, g["done"][t]) and np.array_equal(tout, g["time_out"][t])
        assert np.allclose(obs, g["obs"][t], atol=1e-6) and np.allclose(rew, g["rew"][t], atol=1e-6)
        q, v, _ = b.get_state()
        assert np.allclose(q, g["qpos"][t], atol=1e-6) and np.allclose(v, g["qvel"][t], atol=2e-5)
        assert np.array_equal([b.get(i, "ncon")[0] for i in range(n)], g["ncon"][t])


def test_config1_single_env(oracle_model):
    g = np.load(os.path.join(GOLD, "config1_single_env_1000.npz"))
    assert g["actions"].shape == (1000, 18) and np.abs(g["actions"]).max() <= 1
    import torch
    gen = torch.Generator().manual_seed(0)       # BASELINE.json configs[0] action source
    assert np.array_equal((torch.rand(1000, 18, generator=gen) * 2 - 1).numpy(), g["actions"])
    _replay(oracle_model, g, 1, 0, 300)


def test_batch16(oracle_model):
    g = np.load(os.path.join(GOLD, "batch16_60_steps.npz"))
    _replay(oracle_model, g, 16, 1, 60)
    assert g["done"].sum() == 6 and g["time_out"].sum() == 6
