#!/usr/bin/env python
"""Short PPO training run on the GPU env; prints reward / episode length / throughput per iteration (learning check)."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from envs.helpers import class_to_dict  # noqa: E402
from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO  # noqa: E402
from envs.nightmare_v3_env import NightmareV3Env  # noqa: E402
from rsl_rl.runners import OnPolicyRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
cfg, tc = NightmareV3Config(), NightmareV3ConfigPPO()
cfg.env.num_envs = a.envs
cfg.viewer.render = cfg.viewer.record_states = False
env = NightmareV3Env(cfg, log_dir=tempfile.mkdtemp(), seed=tc.seed)
torch.manual_seed(tc.seed)
runner = OnPolicyRunner(env, class_to_dict(tc), log_dir=None, device="cuda:0")
for it in range(a.iters):
    runner.learn(num_learning_iterations=1, init_at_random_ep_len=(it == 0))
    L = runner.last_log
    print(f"it {it:3d} fps {L['fps']:9d} coll {L['collection_time']:.3f}s learn {L['learn_time']:.3f}s rew {L['mean_reward']} len {L['mean_episode_length']} "
          f"vloss {L['value_loss']:.4f} sloss {L['surrogate_loss']:.4f} lr {L['learning_rate']:.2e} std {L['mean_noise_std']:.3f} "
          f"track {L['episode'].get('rew_tracking_lin_vel', float('nan')):.4f}", flush=True)
