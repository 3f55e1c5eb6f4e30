#!/usr/bin/env python
"""Where does the end-to-end (host buffers in/out) step time go?  Times variants of one 4096-env step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
cfg = NightmareV3Config(); cfg.env.num_envs = E; cfg.viewer.render = cfg.viewer.record_states = False
env = NightmareV3Env(cfg, seed=1, device=dev); env.reset()
env.episode_length_buf = torch.randint(0, 1250, (E,), device=dev)
h_act = torch.randn(16, E, 18).pin_memory()
d_act = h_act.to(dev)
h_obs, h_rew, h_done = torch.empty(E, 66).pin_memory(), torch.empty(E).pin_memory(), torch.empty(E, dtype=torch.int64).pin_memory()
K = 200
def timeit(name, fn):
    for i in range(5): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K): fn(i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K * 1e6
    print(f"{name:55s} {dt:8.1f} us/step  {E / dt:7.2f} M env-steps/s")
b = env._batch
cnt = [1000]
def kern_only(i):
    cnt[0] += 1; b.step(d_act[i % 16], cnt[0])
def kern_sync(i):
    cnt[0] += 1; b.step(d_act[i % 16], cnt[0]); torch.cuda.current_stream().synchronize()
def py_path(i):
    obs, _, rew, done, _ = env.step(h_act[i % 16].to(dev, non_blocking=True))
    h_obs.copy_(obs, non_blocking=True); h_rew.copy_(rew, non_blocking=True); h_done.copy_(done, non_blocking=True)
    torch.cuda.current_stream().synchronize()
def host_api(i):
    env.step_host(h_act[i % 16], h_obs, h_rew, h_done)
def raw_host(i):
    cnt[0] += 1; b.step_host(h_act[i % 16], cnt[0], h_obs, h_rew, h_done)
def copies_only(i):
    d_act[i % 16].copy_(h_act[i % 16], non_blocking=True)
    h_obs.copy_(b.obs, non_blocking=True); h_rew.copy_(b.rew, non_blocking=True); h_done.copy_(b.done, non_blocking=True)
    torch.cuda.current_stream().synchronize()
def d2h_obs_only(i):
    h_obs.copy_(b.obs, non_blocking=True); torch.cuda.current_stream().synchronize()
for rep in range(2):
    timeit("kernel only, async back-to-back", kern_only)
    timeit("kernel + stream sync per step", kern_sync)
    timeit("copies only (H2D act, D2H obs/rew/done) + sync", copies_only)
    timeit("D2H obs only + sync", d2h_obs_only)
    timeit("python path: env.step + torch copies + sync", py_path)
    timeit("env.step_host (public API)", host_api)
    timeit("Batch.step_host (raw C ABI)", raw_host)
