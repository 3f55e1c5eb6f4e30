#!/usr/bin/env python
"""Headless evaluation of a trained policy on the B200 environment step.

The reference's ``play.py`` drives ONE MuJoCo env in an interactive viewer with keyboard commands (``play.py:36-47,81-171``);
a GPU box has neither.  This script keeps what is checkable without a GUI: it finds the checkpoint the way the reference
does (``get_load_path``, ``play.py:66-71``), loads ``model_state_dict`` into ``ActorCritic(66, 66, 18, **policy)``, runs the
policy's MEAN action (``act_inference``; add ``--stochastic`` for ``nn.act`` like ``play.py:122``) on ``--envs`` environments for
``--steps`` control steps and prints reward, episode length and command-tracking statistics.

    python play.py -p logs/nightmare_v3/ [--envs 1024] [--steps 1250] [--command 0.4 0.0 0.0]
"""
import argparse
import os

import torch

from envs.helpers import class_to_dict, get_load_path
from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
from envs.nightmare_v3_env import NightmareV3Env
from nightmare_rl_b200.ppo.policy_kernel import FusedPolicy
from rsl_rl.modules import ActorCritic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-p", "--path", type=str, default="logs/nightmare_v3/", help="log root, run directory or checkpoint file")
    ap.add_argument("-e", "--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=1250)
    ap.add_argument("--stochastic", action="store_true", help="sample actions (nn.act) instead of the mean")
    ap.add_argument("--command", type=float, nargs=3, default=None, metavar=("VX", "VY", "YAW"), help="fixed velocity command for all envs")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    path = args.path
    if os.path.isdir(path):
        path = get_load_path(path) if not any(f.startswith("model_") for f in os.listdir(path)) else get_load_path(os.path.dirname(path.rstrip("/")), load_run=os.path.basename(path.rstrip("/")))
    print(f"Loading model from: {path}")
    cfg, tc = NightmareV3Config(), NightmareV3ConfigPPO()
    cfg.env.num_envs = args.envs
    cfg.viewer.render = False
    cfg.viewer.record_states = False
    if args.command is not None:
        cfg.commands.resampling_time = 1e9                    # keep the given command for the whole run
    env = NightmareV3Env(cfg, seed=args.seed)
    nn = ActorCritic(cfg.env.num_obs, cfg.env.num_obs, cfg.env.num_actions, **class_to_dict(tc)["policy"]).to(env.device)
    loaded = torch.load(path, map_location=env.device, weights_only=False)
    nn.load_state_dict(loaded["model_state_dict"])
    nn.eval()
    policy = FusedPolicy(nn, env.device, seed=args.seed)
    obs, _ = env.reset()
    if args.command is not None:
        env.commands[:] = torch.tensor(args.command, device=env.device)
    rew_sum = torch.zeros(args.envs, device=env.device)
    ep_len = torch.zeros(args.envs, device=env.device)
    done_rew, done_len, n_done = 0.0, 0.0, 0
    track = 0.0
    with torch.inference_mode():
        for t in range(args.steps):
            actions, _, _, _ = policy.act(obs, t + 1, deterministic=not args.stochastic)
            obs, _, rew, dones, _ = env.step(actions)
            if args.command is not None:
                env.commands[:] = torch.tensor(args.command, device=env.device)
            rew_sum += rew
            ep_len += 1
            d = dones > 0
            if t % 50 == 49 or t == args.steps - 1:          # sample the finished episodes now and then (host read)
                nd = int(d.sum())
                if nd:
                    done_rew += float(rew_sum[d].sum()); done_len += float(ep_len[d].sum()); n_done += nd
            rew_sum.masked_fill_(d, 0.0)
            ep_len.masked_fill_(d, 0.0)
            # forward-velocity tracking error against the commanded x velocity (obs[:, 0] = 2 * v_x, obs[:, 9] = 2 * cmd_x)
            track += float(((obs[:, 0] - obs[:, 9]) / 2.0).abs().mean()) if t % 25 == 0 else 0.0
    n_track = len(range(0, args.steps, 25))
    print(f"{args.envs} envs x {args.steps} steps, {'stochastic' if args.stochastic else 'mean'} actions, checkpoint iter {loaded.get('iter')}")
    print(f"mean reward per step {float(env.rew_buf.mean()):.4f} (last step); mean |v_x - cmd_x| {track / n_track:.3f} m/s; "
          f"sampled finished episodes {n_done}" + (f", mean return {done_rew / n_done:.2f}, mean length {done_len / n_done:.1f}" if n_done else ""))


if __name__ == "__main__":
    main()
