#!/usr/bin/env python
"""Fast device-timed throughput check of nm_step at several batch sizes (experiments; bench.py is the contract).

    [NIGHTMARE_B200_LIB=/path/to/variant.so] python tools/quick_bench.py [--sizes 4096,16384,131072] [--steps 50]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config  # noqa: E402
from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env  # noqa: E402

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4096,16384,131072")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--no-flush", action="store_true", help="do not evict L2 between timed steps")
    ap.add_argument("--clean-flush", action="store_true", help="follow the 256 MiB memset by a 256 MiB read of a second buffer: L2 then holds "
                    "clean unrelated lines instead of dirty ones (the step's misses do not have to write anything back first)")
    ap.add_argument("--settle", type=int, default=40, help="untimed steps so that robots are on the ground")
    ap.add_argument("--gait", action="store_true", help="drive the robots with the reference's scripted tripod gait (phase-shifted per env, "
                    "tests/golden/nikengine_gait_targets.npz) instead of N(0,1) actions: walking contacts instead of thrashing")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(7)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush2 = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros((), dtype=torch.int64, device=dev)
    out = []
    for n in [int(x) for x in a.sizes.split(",")]:
        cfg = NightmareV3Config()
        cfg.env.num_envs = n
        cfg.env.model_path = NMB
        cfg.viewer.render = cfg.viewer.record_states = False
        env = NightmareV3Env(cfg, seed=1, device=dev)
        env.reset()
        env.episode_length_buf = torch.randint(0, 1250, (n,), device=dev, generator=gen)
        if a.gait:
            from nightmare_rl_b200.envs.scripted_gait import ScriptedGait
            g = ScriptedGait(os.path.join(ROOT, "tests", "golden", "nikengine_gait_targets.npz"), n, dev, phase_shift=3)
            a.settle = max(a.settle, 330)                        # past the engine's wake-up: everybody is walking when timing starts
            seq = [g.actions().contiguous() for _ in range(a.settle + a.steps)]
            acts = None
        else:
            acts = torch.randn(8, n, 18, device=dev, generator=gen)
            seq = [acts[i % 8] for i in range(a.settle + a.steps)]
        for i in range(a.settle):
            env._batch.step(seq[i], 10 + i)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for i in range(a.steps):
            if not a.no_flush:
                flush.zero_()
                if a.clean_flush:
                    sink.add_(flush2.view(torch.int32)[::1].sum())
            ev[i][0].record()
            env._batch.step(seq[a.settle + i], 1000 + i)
            ev[i][1].record()
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in ev)
        med = ms[len(ms) // 2]
        out.append(f"N={n}: mean {sum(ms) / len(ms) * 1e3:.1f} median {med * 1e3:.1f} us/step  {n / med / 1e3:.1f} M env-steps/s (min {ms[0] * 1e3:.1f} us)")
        done_frac = float(env.reset_buf.float().mean())
        out[-1] += f" resets/step {done_frac:.4f}"
        del env, acts, seq
        torch.cuda.empty_cache()
    print(os.environ.get("NIGHTMARE_B200_LIB", "default lib"), "|", " | ".join(out))


if __name__ == "__main__":
    main()
