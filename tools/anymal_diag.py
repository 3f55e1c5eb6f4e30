"""Diagnostic (GPU box): anymal_c lockstep CUDA vs oracle, dumps the worst env-substeps (inputs and outputs) for offline analysis."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import GenBatch
from oracle import oracle as O

NMB = os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb")
cm = mjcf.CompiledModel.load(NMB)
gm, om, om32 = _lib.GenModel(cm.to_bytes()), O.OracleModel(NMB), O.OracleModel(NMB, variant="f32")
DEV = torch.device("cuda:0")
n, rounds, seed = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 120, 0
rng = np.random.default_rng(seed)
q = np.tile(cm.qpos0, (n, 1)); q[:, 2] = rng.uniform(0.25, 0.7, n)
q[:, 3:7] = rng.normal(size=(n, 4)); q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
q[:, 7:] += rng.uniform(-0.6, 0.6, (n, 12)); v = rng.normal(size=(n, 18)) * 0.5
ob, fb, gb = O.OracleBatch(om, n), O.OracleBatch(om32, n), GenBatch(gm, n, DEV)
ob.set_state(q, v, np.zeros((n, 18)))
ctrl = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
ctrl_d = torch.from_numpy(ctrl).to(DEV)
rec = []
for it in range(rounds):
    q, v, w = ob.get_state()
    q32, v32, w32 = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
    ob.set_state(q32, v32, w32); fb.set_state(q32, v32, w32)
    gb.qpos.copy_(torch.from_numpy(q32)); gb.qvel.copy_(torch.from_numpy(v32)); gb.warm.copy_(torch.from_numpy(w32))
    ob.physics_step(ctrl.astype(np.float64), 1, 8); fb.physics_step(ctrl.astype(np.float64), 1, 8); gb.physics_step(ctrl_d, 1)
    qo, vo, wo = ob.get_state(); qf, vf, wf = fb.get_state()
    vg, wg = gb.qvel.cpu().numpy().astype(np.float64), gb.warm.cpu().numpy().astype(np.float64)
    info = gb.info.cpu().numpy()
    eg = np.abs(vg - vo).max(1) / np.maximum(np.abs(vo).max(1), 1e-3)
    ef = np.abs(vf - vo).max(1) / np.maximum(np.abs(vo).max(1), 1e-3)
    for i in range(n):
        nefc = int(ob.get(i, "nefc")[0]); ncon = ob.get(i, "contact").reshape(-1, 7).shape[0]
        rec.append((it, i, eg[i], ef[i], ncon, nefc, info[i, 0], info[i, 1], info[i, 2], info[i, 3], int(ob.get(i, "solver_niter")[0]) if ob.get(i, "solver_niter").size else -1))
        if eg[i] > 1e-3 and len(rec) < 10**9:
            rec[-1] = rec[-1] + (q32[i].copy(), v32[i].copy(), w32[i].copy(), ctrl[i].copy(), vo[i].copy(), vg[i].copy(), vf[i].copy(), wo[i].copy(), wg[i].copy())
R = np.array([r[:11] for r in rec], dtype=np.float64)
print("n", len(R), "CUDA err pct [50,90,99,99.9,100]", np.percentile(R[:, 2], [50, 90, 99, 99.9, 100]))
print("fp32 oracle err pct", np.percentile(R[:, 3], [50, 90, 99, 99.9, 100]))
bad = [r for r in rec if len(r) > 11]
print("bad (>1e-3):", len(bad))
for r in sorted(bad, key=lambda r: -r[2])[:25]:
    print("it %d env %d err %.2e f32 %.2e ncon %d nefc %d | gpu ncon %d nefc %d iters %d ovf %d | oracle iters %d" % r[:11])
np.savez(os.path.join(ROOT, "gpurun_out", "anymal_diag.npz"), summary=R,
         **{k: np.array([r[11 + j] for r in bad]) for j, k in enumerate(["q", "v", "w", "ctrl", "vo", "vg", "vf", "wo", "wg"])},
         bad_meta=np.array([r[:11] for r in bad]))
