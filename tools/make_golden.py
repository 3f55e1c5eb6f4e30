"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

The reference cannot produce vectors itself (its physics lives in MuJoCo 3.1.2, which is not
installable here), so these fixtures pin the ORACLE: any later change to oracle/ or to
the model compiler that alters results is caught by tests/test_golden.py, and the CUDA path is compared
against the same numbers on the GPU box (where the oracle is rebuilt from source as well).

    python tools/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nightmare_rl_b200.envcfg import build_envcfg            # noqa: E402
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config   # noqa: E402
from oracle import oracle as O                               # noqa: E402

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")
OUT = os.path.join(ROOT, "tests", "golden")


def config1_actions(steps=1000):
    """BASELINE.json configs[0]: a_t ~ U(-1,1)^18 from torch.Generator().manual_seed(0)."""
    g = torch.Generator().manual_seed(0)
    return (torch.rand(steps, 18, generator=g) * 2 - 1).numpy().astype(np.float32)


def rollout(n, actions, seed, ep0=None):
    cfg = NightmareV3Config()
    cfg.env.num_envs = n
    om = O.OracleModel(NMB)
    b = O.OracleBatch(om, n, seed=seed, envcfg=build_envcfg(cfg, 0.008))
    b.env_reset_idx(np.arange(n))
    if ep0 is not None:
        b.env_set("ep_len", ep0)
    rec = dict(obs=[], rew=[], done=[], qpos=[], qvel=[], sens=[], ncon=[], time_out=[], commands=[])
    for a in actions:
        obs, rew, done, tout, _, _ = b.env_step(a.reshape(n, -1))
        q, v, _ = b.get_state()
        rec["obs"].append(obs); rec["rew"].append(rew); rec["done"].append(done); rec["time_out"].append(tout)
        rec["qpos"].append(q.astype(np.float32)); rec["qvel"].append(v.astype(np.float32))
        rec["sens"].append(np.array([b.get(i, "sensordata") for i in range(n)], dtype=np.float32))
        rec["ncon"].append(np.array([b.get(i, "ncon")[0] for i in range(n)], dtype=np.int32))
        rec["commands"].append(b.env_get("commands").astype(np.float32))
    return {k: np.array(v) for k, v in rec.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    a1 = config1_actions()
    r1 = rollout(1, a1[:, None, :], seed=0)
    np.savez_compressed(os.path.join(OUT, "config1_single_env_1000.npz"), actions=a1, **r1)
    rng = np.random.default_rng(1)
    n, steps = 16, 60
    a2 = rng.normal(size=(steps, n, 18)).astype(np.float32)
    ep0 = np.array([0, 100, 620, 621, 1240, 1245, 1249, 1250, 5, 50, 500, 623, 624, 1100, 1200, 1248], dtype=np.float64)
    r2 = rollout(n, a2, seed=1, ep0=ep0)
    np.savez_compressed(os.path.join(OUT, "batch16_60_steps.npz"), actions=a2, ep0=ep0, **r2)
    print("config1: dones", int(r1["done"].sum()), "ncon max", int(r1["ncon"].max()), "| batch16: dones", int(r2["done"].sum()),
          "time_outs", int(r2["time_out"].sum()))


if __name__ == "__main__":
    main()
