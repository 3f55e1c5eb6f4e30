"""Distribution of Newton iterations / contacts in the bench's anymal_c population, and launch time against batch size."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import GenBatch
dev = torch.device("cuda:0")
cm = mjcf.CompiledModel.load(os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb"))
gm = _lib.GenModel(cm.to_bytes())
for E in (256, 1184, 4096, 16384):
    gb = GenBatch(gm, E, dev)
    gen = torch.Generator(device=dev).manual_seed(4321)
    q0 = gb.qpos[0].clone()
    pool = (torch.rand(16, E, 12, device=dev, generator=gen) - 0.5) * 0.7
    def one(i, n=4):
        gb.physics_step(pool[i % 16], n)
        fallen = (gb.qpos[:, 2] < 0.3) | (1.0 - 2.0 * (gb.qpos[:, 4] ** 2 + gb.qpos[:, 5] ** 2) < 0.5)
        gb.qpos[fallen] = q0; gb.qvel[fallen] = 0.0
    for i in range(300):
        one(i)
    its = []
    for i in range(40):
        one(i, 1)
        its.append(gb.info.cpu().numpy().copy())
    its = np.concatenate(its)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        gb.physics_step(pool[i % 16], 4)
    e1.record(); torch.cuda.synchronize()
    print(f"E {E}: {e0.elapsed_time(e1) / 20:.3f} ms per 4-substep launch | iterations pct [50,90,99,99.9,100] {np.percentile(its[:, 2], [50, 90, 99, 99.9, 100])} mean {its[:, 2].mean():.2f} | "
          f"ncon pct [50,99,100] {np.percentile(its[:, 0], [50, 99, 100])} nefc max {its[:, 1].max()} overflow {its[:, 3].sum()}")
    del gb
