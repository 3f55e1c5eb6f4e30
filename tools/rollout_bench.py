#!/usr/bin/env python
"""BASELINE configs[2]: N envs with domain randomisation + PPO rollout (policy forward on tensor cores), rollout only."""
import argparse, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
from envs.nightmare_v3_env import NightmareV3Env
from nightmare_rl_b200.ppo import PPO, ActorCritic
ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=16384); ap.add_argument("--steps", type=int, default=80); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg, tc = NightmareV3Config(), NightmareV3ConfigPPO()
cfg.env.num_envs = a.envs; cfg.viewer.render = cfg.viewer.record_states = False
env = NightmareV3Env(cfg, log_dir=tempfile.mkdtemp(), seed=tc.seed)
env.set_domain_randomization(friction=(0.5, 1.25), kv=(0.8, 1.2), base_mass=(-0.3, 0.3), resample_on_reset=True)
env.reset()
env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=1250)
torch.manual_seed(tc.seed)
ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
alg = PPO(ac, gamma=0.99, device="cuda:0", fused_rollout=True, graph_update=False, seed=1)
alg.init_storage(a.envs, a.steps, [66], [None], [18])
alg.attach_episode_stats(torch.zeros(a.envs, device=dev), torch.zeros(a.envs, device=dev), torch.zeros(100, device=dev), torch.zeros(100, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
assert alg.prepare_fast_rollout(env, torch.zeros(32, device=dev))
best = 1e9
for rep in range(a.reps + 1):
    alg.storage.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(a.steps):
        alg.fast_rollout_step()
    e1.record(); torch.cuda.synchronize()
    if rep:
        best = min(best, e0.elapsed_time(e1))
print(f"{a.envs} envs, DR on, engine {alg.fused.engine}: {a.steps}-step rollout {best:.2f} ms -> {a.envs * a.steps / best / 1e3:.1f} M env-steps/s ({best / a.steps * 1e3:.0f} us per step)")
