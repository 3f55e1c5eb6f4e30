"""Profiling workload for nm_generic_step_kernel (BASELINE configs[3]): the bench's anymal_c population brought to its steady
regime with 300 env steps, then 3 more calls.  Under ncu: -k regex:nm_generic --launch-skip 600 --launch-count 2."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nightmare_rl_b200 import _lib, mjcf
from nightmare_rl_b200.batch import GenBatch

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
cm = mjcf.CompiledModel.load(os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb"))
gb = GenBatch(_lib.GenModel(cm.to_bytes()), E, dev)
gen = torch.Generator(device=dev).manual_seed(4321)
q0 = gb.qpos[0].clone()
pool = (torch.rand(16, E, 12, device=dev, generator=gen) - 0.5) * 0.7
for i in range(303):
    gb.physics_step(pool[i % 16], 4)
    fallen = (gb.qpos[:, 2] < 0.3) | (1.0 - 2.0 * (gb.qpos[:, 4] ** 2 + gb.qpos[:, 5] ** 2) < 0.5)
    gb.qpos[fallen] = q0
    gb.qvel[fallen] = 0.0
torch.cuda.synchronize()
print("ok", float(gb.info[:, 0].float().mean()), float(gb.info[:, 2].float().mean()))
