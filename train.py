#!/usr/bin/env python
"""PPO training of the Nightmare-v3 hexapod on the B200 environment step.

Same command line and control flow as the reference's ``train.py`` (flags ``-r -v -n -e -p``, log directory
``logs/nightmare_v3/<timestamp>/``, resume via ``get_load_path``, ``learn(max_iterations, init_at_random_ep_len=True)``),
plus what a multi-GPU box needs: launched under ``torchrun`` every rank owns ``--envs`` environments on its own GPU and
gradients are all-reduced over NCCL (``--iterations`` bounds the run, default: the config's max_iterations).

    python train.py -e 4096
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 train.py -e 16384
"""
import argparse
import datetime
import os

import torch
import torch.distributed as dist

from envs.helpers import class_to_dict, get_load_path
from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
from envs.nightmare_v3_env import NightmareV3Env
from rsl_rl.runners import OnPolicyRunner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-r", "--resume", action="store_true", default=False, help="resume from the newest checkpoint")
    ap.add_argument("-v", "--render", action="store_true", default=False, help="accepted for compatibility; there is no viewer on a GPU box")
    ap.add_argument("-n", "--num_threads", type=int, default=1, help="accepted for compatibility; the GPU step needs no host threads")
    ap.add_argument("-e", "--envs", type=int, default=2048, dest="num_envs", help="environments PER GPU")
    ap.add_argument("-p", "--resume_path", type=str, default=None, help="log root (or run) to resume from")
    ap.add_argument("--iterations", type=int, default=None, help="number of PPO iterations (default: train_cfg.runner.max_iterations)")
    ap.add_argument("--log_root", type=str, default="logs/nightmare_v3/")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    log_dir = f"{args.log_root}{datetime.datetime.now()}/"
    if rank == 0:
        print(f"Resume: {args.resume}, Render: {args.render}, Num threads: {args.num_threads}, ranks: {world}")
        print(f"Logging to {log_dir}")

    cfg = NightmareV3Config()
    train_cfg = NightmareV3ConfigPPO()
    cfg.viewer.render = args.render
    cfg.env.num_envs = args.num_envs
    cfg.rl_device = f"cuda:{local}"
    if rank != 0:
        cfg.viewer.record_states = False
    train_cfg.runner.resume = args.resume
    train_cfg_dict = class_to_dict(train_cfg)

    env = NightmareV3Env(cfg, log_dir=log_dir, num_threads=args.num_threads, seed=train_cfg.seed, env_offset=rank * args.num_envs)
    torch.manual_seed(train_cfg.seed)
    runner = OnPolicyRunner(env, train_cfg_dict, log_dir=log_dir, device=cfg.rl_device)

    if train_cfg.runner.resume:
        root = args.resume_path if args.resume_path is not None else args.log_root
        path = get_load_path(root, load_run=train_cfg.runner.load_run, checkpoint=train_cfg.runner.checkpoint)
        if rank == 0:
            print(f"Loading model from: {path}")
        runner.load(path)

    iters = args.iterations if args.iterations is not None else train_cfg.runner.max_iterations
    runner.learn(num_learning_iterations=iters, init_at_random_ep_len=True)
    if world > 1:
        runner.alg.release_graph()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
