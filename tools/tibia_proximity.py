#!/usr/bin/env python
"""How often could the tibia-tibia convex-convex pairs of mjmodel.xml (contype=2/conaffinity=3, 15 pairs) matter?

Neither the oracle nor the CUDA step collides tibias with each other.  MuJoCo would run its convex narrow
phase on the pairs whose bounding spheres overlap and create a contact only if the hulls intersect.  This tool replays the
bench workload (NightmareV3Env semantics, N(0,1) actions, random episode phases) on the CPU oracle and measures, for all 15
pairs, the minimum distance between the two hulls' vertex sets and whether any vertex of one hull lies inside the other
(scipy Delaunay).  A pair that never comes closer than a few millimetres can not produce a contact in MuJoCo either, i.e.
the omission does not change those trajectories.

    python tools/tibia_proximity.py [--envs 128] [--steps 300]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scipy.spatial import Delaunay, cKDTree  # noqa: E402

from nightmare_rl_b200 import mjcf  # noqa: E402
from nightmare_rl_b200.envcfg import build_envcfg  # noqa: E402
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config  # noqa: E402
from oracle import oracle as O  # noqa: E402

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=128)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--every", type=int, default=3)
    ap.add_argument("--gait", action="store_true", help="drive ONE env with the reference's scripted tripod gait "
                    "(tests/golden/nikengine_gait_targets.npz) instead of random actions")
    a = ap.parse_args()
    gait = None
    if a.gait:
        z = np.load(os.path.join(ROOT, "tests", "golden", "nikengine_gait_targets.npz"))
        gait = ((z["targets"] + np.array([0.0, np.pi / 5, 0.0] * 6)) / 0.2).astype(np.float32)
        a.envs, a.steps = 1, len(gait)
    cm = mjcf.CompiledModel.load(NMB)
    A = cm.arrays
    hv = A["hull_vert"].reshape(-1, 3).astype(np.float64)
    tib_geoms = [g for g in range(cm.ngeom) if A["geom_hull_num"][g] > 0 and A["geom_body"][g] > 1]
    hulls = [hv[A["geom_hull_adr"][g]: A["geom_hull_adr"][g] + A["geom_hull_num"][g]] for g in tib_geoms]
    bodies = [int(A["geom_body"][g]) for g in tib_geoms]
    gpos = [A["geom_pos"].reshape(-1, 3)[g] for g in tib_geoms]
    gquat = [A["geom_quat"].reshape(-1, 4)[g] for g in tib_geoms]

    def q2m(q):
        w, x, y, z = q
        return np.array([[w*w+x*x-y*y-z*z, 2*(x*y-w*z), 2*(x*z+w*y)], [2*(x*y+w*z), w*w-x*x+y*y-z*z, 2*(y*z-w*x)], [2*(x*z-w*y), 2*(y*z+w*x), w*w-x*x-y*y+z*z]])
    gmat = [q2m(q) for q in gquat]
    cfg = NightmareV3Config()
    cfg.env.num_envs = a.envs
    om = O.OracleModel(NMB)
    ob = O.OracleBatch(om, a.envs, seed=1, envcfg=build_envcfg(cfg, 0.008))
    ob.env_reset_idx(np.arange(a.envs))
    rng = np.random.default_rng(1)
    if gait is None:
        ob.env_set("ep_len", rng.integers(0, 1250, a.envs).astype(np.float64))
    mind, inside, samples, fallen = [], 0, 0, 0
    for t in range(a.steps):
        ob.env_step(gait[t][None] if gait is not None else rng.normal(size=(a.envs, 18)).astype(np.float32), 8)
        if t % a.every:
            continue
        ob.forward(None, 8)
        for i in range(a.envs):
            xpos = ob.get(i, "xpos").reshape(-1, 3)
            xmat = ob.get(i, "xmat").reshape(-1, 3, 3)
            W = []
            for k, b in enumerate(bodies):                      # hull vertices are stored in the BODY frame by the compiler
                W.append(hulls[k] @ xmat[b].T + xpos[b])
            best = np.inf
            hit = False
            for p in range(6):
                for q in range(p + 1, 6):
                    if np.linalg.norm(W[p].mean(0) - W[q].mean(0)) > 0.25:        # bounding spheres (rbound 0.115) apart
                        continue
                    d = cKDTree(W[p]).query(W[q])[0].min()
                    best = min(best, d)
                    if d < 0.02:
                        hit |= bool((Delaunay(W[p]).find_simplex(W[q]) >= 0).any() or (Delaunay(W[q]).find_simplex(W[p]) >= 0).any())
            mind.append(best)
            inside += int(hit)
            samples += 1
    mind = np.array(mind)
    fin = mind[np.isfinite(mind)]
    print(f"{samples} env-states sampled from {a.envs} envs x {a.steps} {'scripted-gait' if gait is not None else 'random-action'} steps")
    print(f"closest approach of any tibia pair (vertex-to-vertex): min {fin.min() * 1e3:.1f} mm, 1st percentile {np.percentile(fin, 1) * 1e3:.1f} mm, median {np.median(fin) * 1e3:.1f} mm")
    print(f"states with a tibia pair closer than 5 mm: {(fin < 0.005).sum()} ({100.0 * (fin < 0.005).sum() / samples:.3f} %), closer than 20 mm: {(fin < 0.02).sum()}")
    print(f"states with interpenetrating tibia hulls (vertex-in-hull test): {inside} ({100.0 * inside / samples:.3f} %)")


if __name__ == "__main__":
    main()
