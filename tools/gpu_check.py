"""Ad-hoc GPU-vs-oracle comparison (development aid; the real checks live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import Batch
from nightmare_rl_b200.envcfg import build_envcfg
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from oracle import oracle as O

NMB = "models/nightmare_v3/mjmodel.nmb"
cm = mjcf.CompiledModel.load(NMB)
dm = _lib.Model(cm.to_bytes())
om = O.OracleModel(NMB)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)

def relerr(a, b):
    return np.abs(a - b).max() / max(1e-12, np.abs(b).max())

# ---- 1. contact-free single substep from random states
N = 256
qpos = np.tile(cm.qpos0, (N, 1)); qpos[:, 2] = 1.0
qpos[:, 3:7] = rng.normal(size=(N, 4)); qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
qpos[:, 7:] = rng.uniform(-0.8, 0.8, (N, 18))
qvel = rng.normal(size=(N, 24)) * np.r_[np.ones(3) * 0.5, np.ones(3) * 2, np.ones(18) * 3]
ctrl = rng.uniform(-10, 10, (N, 18))
qpos32, qvel32, ctrl32 = qpos.astype(np.float32), qvel.astype(np.float32), ctrl.astype(np.float32)
ob = O.OracleBatch(om, N)
ob.set_state(qpos32, qvel32, np.zeros((N, 24)))
ob.physics_step(ctrl32, 1, 8)
oq, ov, ow = ob.get_state()
gb = Batch(dm, N, dev, debug=True)
gb.qpos.copy_(torch.from_numpy(qpos32)); gb.qvel.copy_(torch.from_numpy(qvel32))
gb.physics_step(torch.from_numpy(ctrl32), 1)
torch.cuda.synchronize()
gq, gv, gw = gb.qpos.cpu().numpy(), gb.qvel.cpu().numpy(), gb.warm.cpu().numpy()
print("free: qpos rel", relerr(gq, oq), "qvel rel", relerr(gv, ov), "warm rel", relerr(gw, ow))
dbg = gb.debug.cpu().numpy()
oqs = np.array([ob.get(i, "qacc_smooth") for i in range(N)])
gqs = np.concatenate([dbg[:, 96:102], dbg[:, 102:120]], axis=1)
print("      qacc_smooth rel", relerr(gqs, oqs), "per-dof max abs", np.abs(gqs - oqs).max(0)[:8])

# ---- 2. drop test with random ctrl, lockstep (re-sync GPU state from oracle each substep)
N = 512
ob = O.OracleBatch(om, N); gb = Batch(dm, N, dev, debug=True)
qpos = np.tile(cm.qpos0, (N, 1)); qpos[:, 7:] += rng.uniform(-0.3, 0.3, (N, 18)); qpos[:, 2] = rng.uniform(0.02, 0.16, N)
qpos[:, 3:7] += rng.normal(size=(N, 4)) * 0.1; qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
ob.set_state(qpos.astype(np.float32), np.zeros((N, 24)), np.zeros((N, 24)))
worst = dict(qpos=0, qvel=0, warm=0, sens=0); mism = 0; tot = 0; ncmax = 0
for t in range(120):
    ctrl = rng.uniform(-8, 8, (N, 18)).astype(np.float32) if t % 4 == 0 else ctrl
    q, v, w = ob.get_state()
    q32, v32, w32 = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
    ob.set_state(q32, v32, w32)          # both sides start from the same fp32 state
    gb.qpos.copy_(torch.from_numpy(q32)); gb.qvel.copy_(torch.from_numpy(v32)); gb.warm.copy_(torch.from_numpy(w32))
    ob.physics_step(ctrl, 1, 8)
    gb.physics_step(torch.from_numpy(ctrl), 1)
    torch.cuda.synchronize()
    oq, ov, ow = ob.get_state()
    gq, gv, gw = gb.qpos.cpu().numpy(), gb.qvel.cpu().numpy(), gb.warm.cpu().numpy()
    osens = np.array([ob.get(i, "sensordata") for i in range(N)]); gsens = gb.sensordata.cpu().numpy()
    oncon = np.array([ob.get(i, "ncon")[0] for i in range(N)]); gncon = gb.debug[:, 0].cpu().numpy()
    bad = oncon != gncon
    mism += bad.sum(); tot += N; ncmax = max(ncmax, oncon.max())
    ok = ~bad
    dv = np.abs(gv - ov)[ok].max() / max(1e-9, np.abs(ov).max())
    worst["qpos"] = max(worst["qpos"], np.abs(gq - oq)[ok].max() / np.abs(oq).max())
    worst["qvel"] = max(worst["qvel"], dv)
    worst["warm"] = max(worst["warm"], np.abs(gw - ow)[ok].max() / max(1e-9, np.abs(ow).max()))
    worst["sens"] = max(worst["sens"], np.abs(gsens - osens)[ok].max() / max(1.0, np.abs(osens).max()))
    if t % 20 == 0 or dv > 1e-3:
        i = int(np.abs(gv - ov).max(1).argmax())
        print(t, "ncon mean", oncon.mean(), "mismatch", bad.sum(), "qvel rel", dv, "worst env", i, oncon[i], gncon[i])
print("lockstep worst rel errs", worst, "ncon mismatches", mism, "/", tot, "max ncon", ncmax)

# ---- 3. timing
N = 4096
cfg = NightmareV3Config(); ec = build_envcfg(cfg, 0.008)
gb = Batch(dm, N, dev, seed=1, envcfg=ec)
act = torch.randn(N, 18, device=dev)
for i in range(20): gb.step(act, i + 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200): gb.step(act, 21 + i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print(f"env step N={N}: {ms*1e3:.1f} us/step -> {N/ms*1e3:.3e} env-steps/s; dones {gb.done.sum().item()} rew mean {gb.rew.mean().item():.4f}")
