#!/usr/bin/env python
"""Raw physics throughput, the B200 counterpart of the reference's `simple_test.py`.

The reference script builds N `MjData`, sets `model.opt.timestep = 0.0025`, and times `mj_step(model, data[i], decimation)`
with zero controls over `num_threads` Python threads (reference simple_test.py:21-47), printing SUBSTEPS per second.  Here the
same workload is one `nm_physics_step` launch per outer step for all N robots; the flags keep their names (`-t` is accepted
and ignored: there are no host threads to configure) and the printed unit is the same, so numbers can be put side by side.

    python simple_test.py -e 4096 -s 200 -d 4
"""
import argparse
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("-d", "--decimation", type=int, default=4, help="physics substeps per outer step")
    ap.add_argument("-e", "--env_num", type=int, default=2048, help="number of robots")
    ap.add_argument("-s", "--num_steps", type=int, default=10, help="outer steps to time")
    ap.add_argument("-t", "--num_threads", type=int, default=12, help="ignored (kept for command-line compatibility)")
    ap.add_argument("--timestep", type=float, default=0.0025, help="model.opt.timestep (the reference script overrides it to 0.0025)")
    ap.add_argument("--model", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "models", "nightmare_v3", "mjmodel.nmb"))
    args = ap.parse_args()

    from nightmare_rl_b200 import _lib, mjcf
    from nightmare_rl_b200.batch import Batch

    if not torch.cuda.is_available():
        raise _lib.NightmareLibError("simple_test.py needs a CUDA device: the physics step has no CPU path")
    compiled = mjcf.load_model(args.model)
    compiled.arrays["opt_real"][0] = args.timestep
    device = torch.device("cuda", torch.cuda.current_device())
    batch = Batch(_lib.Model(compiled.to_bytes()), args.env_num, device)
    ctrl = torch.zeros(args.env_num, 18, device=device)

    for _ in range(3):                                       # warm-up: module load, first-touch of the state buffers
        batch.physics_step(ctrl, args.decimation)
    torch.cuda.synchronize()
    begin = time.time()
    for _ in range(args.num_steps):
        batch.physics_step(ctrl, args.decimation)
    torch.cuda.synchronize()
    elapsed = time.time() - begin
    print(args.env_num * args.num_steps * args.decimation / elapsed, "steps per second")


if __name__ == "__main__":
    main()
