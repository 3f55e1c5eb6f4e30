import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import Batch
from oracle import oracle as O
NMB = "models/nightmare_v3/mjmodel.nmb"
cm = mjcf.CompiledModel.load(NMB); dm = _lib.Model(cm.to_bytes()); om = O.OracleModel(NMB)
dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
N = 512
ob = O.OracleBatch(om, N); gb = Batch(dm, N, dev, debug=True)
qpos = np.tile(cm.qpos0, (N, 1)); qpos[:, 7:] += rng.uniform(-0.3, 0.3, (N, 18)); qpos[:, 2] = rng.uniform(0.02, 0.16, N)
qpos[:, 3:7] += rng.normal(size=(N, 4)) * 0.1; qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
ob.set_state(qpos.astype(np.float32), np.zeros((N, 24)), np.zeros((N, 24)))
errs = []; flagmis = []
for t in range(60):
    ctrl = rng.uniform(-8, 8, (N, 18)).astype(np.float32) if t % 4 == 0 else ctrl
    q, v, w = ob.get_state(); q32, v32, w32 = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
    ob.set_state(q32, v32, w32)
    gb.qpos.copy_(torch.from_numpy(q32)); gb.qvel.copy_(torch.from_numpy(v32)); gb.warm.copy_(torch.from_numpy(w32))
    ob.physics_step(ctrl, 1, 8); gb.physics_step(torch.from_numpy(ctrl), 1); torch.cuda.synchronize()
    oq, ov, ow = ob.get_state(); gv = gb.qvel.cpu().numpy(); dbg = gb.debug.cpu().numpy()
    oflag = np.array([ob.get(i, "solver_niter") for i in range(N)])   # pgs, noslip, warm
    gflag = dbg[:, [1, 2, 3]]
    oacc = np.array([ob.get(i, "qacc") for i in range(N)]); gacc = np.concatenate([dbg[:, 128:134], dbg[:, 134:152]], 1)
    oqs = np.array([ob.get(i, "qacc_smooth") for i in range(N)]); gqs = np.concatenate([dbg[:, 96:102], dbg[:, 102:120]], 1)
    e = np.abs(gv - ov).max(1) / np.abs(ov).max(1).clip(1e-3)
    ea = np.abs(gacc - oacc).max(1) / np.abs(oacc).max(1).clip(1e-3)
    es = np.abs(gqs - oqs).max(1) / np.abs(oqs).max(1).clip(1e-3)
    fm = (oflag != gflag).any(1)
    errs.append(e); flagmis.append(fm)
    if t < 6 or t % 10 == 0:
        i = int(e.argmax())
        print(f"t={t} qvel relerr per-env: median {np.median(e):.2e} p99 {np.percentile(e,99):.2e} max {e.max():.2e} | qacc max {ea.max():.2e} smooth max {es.max():.2e} | flag mism {fm.sum()} | worst env {i}: oracle flags {oflag[i]} gpu flags {gflag[i]} ncon {dbg[i,0]} |ov| {np.abs(ov[i]).max():.2f} fmis {fm[i]}")
errs = np.array(errs); flagmis = np.array(flagmis)
print("with flag match: max", errs[~flagmis].max(), "p99.9", np.percentile(errs[~flagmis], 99.9), " | with flag mismatch: n", flagmis.sum(), "max", errs[flagmis].max() if flagmis.any() else 0)
