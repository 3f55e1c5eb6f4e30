"""Would grouping anymal_c environments by their Newton iteration count cut the lockstep barrier wait?  Upper bound: permute the
state so that the natural order is sorted by last step's iteration count, then time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import GenBatch
dev = torch.device("cuda:0")
cm = mjcf.CompiledModel.load(os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb"))
E = 4096
gb = GenBatch(_lib.GenModel(cm.to_bytes()), E, dev)
gen = torch.Generator(device=dev).manual_seed(4321)
q0 = gb.qpos[0].clone()
pool = (torch.rand(16, E, 12, device=dev, generator=gen) - 0.5) * 0.7
def one(i, ctrl):
    gb.physics_step(ctrl, 4)
    fallen = (gb.qpos[:, 2] < 0.3) | (1.0 - 2.0 * (gb.qpos[:, 4] ** 2 + gb.qpos[:, 5] ** 2) < 0.5)
    gb.qpos[fallen] = q0; gb.qvel[fallen] = 0.0
for i in range(300):
    one(i, pool[i % 16])
def timed(sort):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(40):
        ctrl = pool[i % 16]
        if sort:
            order = torch.argsort(gb.info[:, 2].to(torch.int64) * 64 + gb.info[:, 0].to(torch.int64), stable=True)
            for t in (gb.qpos, gb.qvel, gb.warm):
                t.copy_(t[order].clone())
            pool.copy_(pool[:, order].clone())
            ctrl = pool[i % 16]
        e0.record(); gb.physics_step(ctrl, 4); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        fallen = (gb.qpos[:, 2] < 0.3) | (1.0 - 2.0 * (gb.qpos[:, 4] ** 2 + gb.qpos[:, 5] ** 2) < 0.5)
        gb.qpos[fallen] = q0; gb.qvel[fallen] = 0.0
    return tot / 40
a = timed(False); b = timed(True); c = timed(False)
print(f"physics launch: natural order {a:.3f} ms, sorted by (Newton iterations, contacts) of the previous call {b:.3f} ms, natural again {c:.3f} ms")
