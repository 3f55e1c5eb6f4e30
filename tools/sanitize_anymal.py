"""Small anymal_c run for compute-sanitizer: both capacity tiers (tumbling robots overflow the first), and a crossed-legs hexapod
batch that enters the cold MPR path of the step kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import GenBatch, Batch
dev = torch.device("cuda:0")
cm = mjcf.CompiledModel.load(os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb"))
n = 203
rng = np.random.default_rng(0)
q = np.tile(cm.qpos0, (n, 1)); q[:, 2] = rng.uniform(0.25, 0.7, n)
q[:, 3:7] = rng.normal(size=(n, 4)); q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
q[:, 7:] += rng.uniform(-0.6, 0.6, (n, 12))
gb = GenBatch(_lib.GenModel(cm.to_bytes()), n, dev)
gb.qpos.copy_(torch.from_numpy(q.astype(np.float32)))
ctrl = torch.from_numpy(rng.uniform(-1, 1, (n, 12)).astype(np.float32)).to(dev)
for t in range(30):
    gb.physics_step(ctrl, 4)
torch.cuda.synchronize()
info = gb.info.cpu().numpy()
print("anymal ok", n, "max ncon", info[:, 0].max(), "max nefc", info[:, 1].max(), "finite", bool(torch.isfinite(gb.qpos).all()))
# hexapod with crossed legs in the air: tibia-tibia pairs
hm = mjcf.CompiledModel.load(os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb"))
n = 150
q = np.tile(hm.qpos0, (n, 1)); q[:, 2] = 1.0
q[:, 7:] += rng.uniform(-1.0, 1.0, (n, 18))
hb = Batch(_lib.Model(hm.to_bytes()), n, dev, debug=True)
hb.qpos.copy_(torch.from_numpy(q.astype(np.float32)))
c = torch.from_numpy(rng.uniform(-8, 8, (n, 18)).astype(np.float32)).to(dev)
for t in range(20):
    hb.physics_step(c, 2)
torch.cuda.synchronize()
print("hexapod pairs ok", n, "pair contacts seen", int((hb.debug[:, 0] > 0).sum()), "finite", bool(torch.isfinite(hb.qpos).all()))
