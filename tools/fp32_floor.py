"""Arithmetic floor of an fp32 implementation of the step, measured on the CPU.

`oracle/libnm_oracle_f32.so` is the oracle's own source compiled with `real = float` (same dense algorithm,
same order of operations, float constants and float libm).  This tool replays the workloads of the GPU
lockstep suites (tests/test_gpu_physics.py::test_in_contact_lockstep, tests/test_gpu_env.py::_lockstep)
with that build standing where the CUDA kernel stands in the tests: before every substep both sides restart
from the fp64 oracle's state rounded to fp32, so every comparison is a one-step comparison on identical
inputs.  The spread it prints is what *rounding alone* does to this pipeline in fp32; the CUDA kernel's
spread is judged against it (profiles/r02_fp32_floor.md).

    python tools/fp32_floor.py [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")


def per_env_rel(a, b, floor=1e-3):
    return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), floor)


def elem_rel(a, b, floor):
    """element-wise |a-b| / max(|b|, floor), worst element per env"""
    return (np.abs(a - b) / np.maximum(np.abs(b), floor)).max(axis=1)


def pct(x):
    return dict(median=float(np.median(x)), p99=float(np.percentile(x, 99)), p999=float(np.percentile(x, 99.9)), max=float(x.max()))


def physics_lockstep(tumbling, n=512, T=60):
    from nightmare_rl_b200 import mjcf
    cm = mjcf.CompiledModel.load(NMB)
    m64, m32 = O.OracleModel(NMB), O.OracleModel(NMB, variant="f32")
    rng = np.random.default_rng(1 if tumbling else 0)
    a, b = O.OracleBatch(m64, n), O.OracleBatch(m32, n)
    qpos = np.tile(cm.qpos0, (n, 1))
    qvel = np.zeros((n, 24))
    if tumbling:
        qpos[:, 7:] += rng.uniform(-0.6, 0.6, (n, 18))
        qpos[:, 2] = rng.uniform(0.03, 0.22, n)
        qpos[:, 3:7] = rng.normal(size=(n, 4))
        qvel[:, 3:6] = rng.uniform(-3, 3, (n, 3))
        qvel[:, 0:3] = rng.uniform(-0.5, 0.5, (n, 3))
    else:
        qpos[:, 7:] += rng.uniform(-0.3, 0.3, (n, 18))
        qpos[:, 2] = rng.uniform(0.02, 0.16, n)
        qpos[:, 3:7] += rng.normal(size=(n, 4)) * 0.1
    qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
    a.set_state(qpos.astype(np.float32), qvel.astype(np.float32), np.zeros((n, 24)))
    ev, eq, ee, stage = [], [], [], {k: [] for k in ("qacc_smooth", "efc_b", "efc_force", "qacc")}
    ncon_diff = flag_diff = 0
    for t in range(T):
        if t % 4 == 0:
            ctrl = rng.uniform(-8, 8, (n, 18)).astype(np.float32)
        q, v, w = a.get_state()
        q32, v32, w32 = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
        a.set_state(q32, v32, w32)
        b.set_state(q32, v32, w32)
        a.physics_step(ctrl, 1, 8)
        b.physics_step(ctrl, 1, 8)
        oq, ov, _ = a.get_state()
        fq, fv, _ = b.get_state()
        nca = np.array([a.get(i, "ncon")[0] for i in range(n)])
        ncb = np.array([b.get(i, "ncon")[0] for i in range(n)])
        same = nca == ncb
        # same contact set (vertex ids) only: a different support vertex is a tie, not rounding
        for i in np.nonzero(same & (nca > 0))[0]:
            ca, cb_ = a.get(i, "contact").reshape(-1, 7), b.get(i, "contact").reshape(-1, 7)
            if not np.array_equal(ca[:, :3], cb_[:, :3]):
                same[i] = False
        ncon_diff += int((~same).sum())
        fa = np.array([a.get(i, "solver_niter") for i in range(n)])
        fb = np.array([b.get(i, "solver_niter") for i in range(n)])
        flag_diff += int((fa != fb).any(axis=1).sum())
        ev.append(per_env_rel(fv, ov)[same]); eq.append(per_env_rel(fq, oq)[same])
        ee.append(elem_rel(fv, ov, 1e-2)[same])
        for k in stage:
            for i in np.nonzero(same & (nca > 0))[0][:64]:
                x, y = a.get(i, k), b.get(i, k)
                if x.size and x.size == y.size:
                    stage[k].append(np.abs(x - y).max() / max(np.abs(x).max(), 1e-3))
    ev, eq, ee = np.concatenate(ev), np.concatenate(eq), np.concatenate(ee)
    return dict(samples=int(ev.size), contact_set_differs=ncon_diff, solver_flag_differs=flag_diff,
                qvel_per_env_rel=pct(ev), qpos_per_env_rel=pct(eq), qvel_elementwise_rel_floor1e2=pct(ee),
                stages_rel={k: pct(np.array(vv)) for k, vv in stage.items() if vv})


def env_lockstep(n=256, T=50, seed=1, action_scale=1.0):
    from nightmare_rl_b200.envcfg import build_envcfg
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    cfg = NightmareV3Config()
    cfg.env.num_envs = n
    ec = build_envcfg(cfg, 0.008)
    m64, m32 = O.OracleModel(NMB), O.OracleModel(NMB, variant="f32")
    a, b = O.OracleBatch(m64, n, seed=seed, envcfg=ec), O.OracleBatch(m32, n, seed=seed, envcfg=ec)
    a.env_reset_idx(np.arange(n)); b.env_reset_idx(np.arange(n))
    rng = np.random.default_rng(0)
    ep0 = rng.integers(0, 1251, n).astype(np.float64)
    ep0[:8] = [620, 624, 1245, 1249, 1250, 0, 623, 1248]
    a.env_set("ep_len", ep0)
    rng = np.random.default_rng(seed)
    er, eo, flags = [], [], 0
    for t in range(T):
        q, v, w = a.get_state()
        q32, v32, w32 = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
        a.set_state(q32, v32, w32); b.set_state(q32, v32, w32)
        for name in ("actions", "dof_pos", "dof_vel", "commands", "episode_sums"):
            val = a.env_get(name).astype(np.float32)
            a.env_set(name, val); b.env_set(name, val)
        b.env_set("ep_len", a.env_get("ep_len"))
        b.env_set("step_counter", [a.env_get("step_counter")])
        act = (rng.normal(size=(n, 18)) * action_scale).astype(np.float32)
        obs, rew, done, *_ = a.env_step(act)
        obs2, rew2, done2, *_ = b.env_step(act)
        ok = done == done2
        flags += int((~ok).sum())
        er.append(np.abs(rew2 - rew)[ok]); eo.append(np.abs(obs2 - obs)[ok].max(axis=1))
    er, eo = np.concatenate(er), np.concatenate(eo)
    return dict(samples=int(er.size), reset_flag_differs=flags, rew_abs=pct(er), obs_abs=pct(eo))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    out = dict(
        what="fp32 build of the oracle vs the fp64 oracle, one-step comparisons from identical fp32 states",
        physics_near_upright=physics_lockstep(False),
        physics_tumbling=physics_lockstep(True),
        env_default=env_lockstep(),
    )
    print(json.dumps(out, indent=1))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
