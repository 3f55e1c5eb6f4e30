"""Diagnostic: free-running GPU vs oracle divergence during the landing/settle sequence of tests/test_gpu_physics.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpu_common as G

cm, dm, om = G.models()
n = 64
rng = np.random.default_rng(3)
ob = G.O.OracleBatch(om, n)
gb = G.Batch(dm, n, G.DEV, debug=True)
target = np.tile(np.array([0, np.pi / 5, 0] * 6), (n, 1)) * -1 + rng.uniform(-0.05, 0.05, (n, 18))
for t in range(160):
    q, _, _ = ob.get_state()
    ctrl = ((target - q[:, 7:]) * 20.0).astype(np.float32)
    ob.physics_step(ctrl, 2, 8)
    gb.physics_step(torch.from_numpy(ctrl), 2)
    torch.cuda.synchronize()
    oq, ov, _ = ob.get_state()
    gq, gv, _ = G.gpu_state(gb)
    eq = G.per_env_rel(gq, oq); ev = G.per_env_rel(gv, ov, floor=0.1)
    ncon = np.array([ob.get(i, "ncon")[0] for i in range(n)])
    gncon = gb.debug[:, 0].cpu().numpy()
    if t % 5 == 0 or t < 20:
        i = int(ev.argmax())
        print(f"t={t:3d} ncon(oracle) min/max {ncon.min()}/{ncon.max()} ncon mismatch envs {(ncon != gncon).sum():2d} | q rel max {eq.max():.2e} | v rel max {ev.max():.2e} (env {i}, |v|max {np.abs(ov[i]).max():.3f}, abs err {np.abs(gv[i]-ov[i]).max():.2e}) median v err {np.median(ev):.2e}")
