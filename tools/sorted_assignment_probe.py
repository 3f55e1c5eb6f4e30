#!/usr/bin/env python
"""Upper bound of what assigning environments to warps by contact load would buy: permute the batch's state so that the
natural env order IS sorted by the number of touching geoms, then time a few steps (contact patterns persist for tens of steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(7)
for n in (16384, 131072):
    cfg = NightmareV3Config(); cfg.env.num_envs = n; cfg.env.model_path = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")
    cfg.viewer.render = cfg.viewer.record_states = False
    env = NightmareV3Env(cfg, seed=1, device=dev); env.reset()
    env.episode_length_buf = torch.randint(0, 1250, (n,), device=dev, generator=gen)
    acts = torch.randn(8, n, 18, device=dev, generator=gen)
    b = env._batch
    for i in range(60):
        b.step(acts[i % 8], 10 + i)
    def timed(k0, reps=12):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            b.step(acts[(k0 + i) % 8], 1000 + k0 + i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    t_unsorted = timed(0)
    key = (b.sensordata > 0).sum(1)
    order = torch.argsort(key, stable=True)
    for name in ("qpos", "qvel", "warm", "actions", "dof_pos", "dof_vel", "commands", "episode_length", "episode_sums", "feet_air_time", "contact_bits"):
        t = getattr(b, name); t.copy_(t[order].clone())
    acts = acts[:, order].contiguous()
    for i in range(3):
        b.step(acts[i % 8], 2000 + i)                      # hull hints re-learned
    t_sorted = timed(3)
    print(f"N={n}: natural order {t_unsorted:.1f} us/step, sorted by touching geoms {t_sorted:.1f} us/step ({100 * (t_unsorted / t_sorted - 1):.1f} % faster); "
          f"key histogram {torch.bincount(key).tolist()}")
    del env, b
    torch.cuda.empty_cache()
