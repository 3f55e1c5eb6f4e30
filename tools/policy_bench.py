#!/usr/bin/env python
"""Device time of the fused policy forward (both engines) at several batch sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from nightmare_rl_b200.ppo import ActorCritic
from nightmare_rl_b200.ppo.policy_kernel import FusedPolicy
dev = torch.device("cuda:0")
ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30]).to(dev)
for eng in ("tc5", "mma"):
    fp = FusedPolicy(ac, dev, seed=1, engine=eng)
    out = []
    for n in (4096, 16384, 131072):
        obs = torch.randn(n, 66, device=dev)
        o = (torch.empty(n, 18, device=dev), torch.empty(n, 18, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))
        for i in range(5): fp.act(obs, i, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50): fp.act(obs, 10 + i, out=o)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        out.append(f"N={n}: {us:.1f} us ({n / us:.1f} M obs/s)")
    print(eng, "|", " | ".join(out))
