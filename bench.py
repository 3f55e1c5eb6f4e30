#!/usr/bin/env python
"""bench.py — hexapod env-steps/s of the batched Nightmare-v3 environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs-per-gpu E]

One "step" = one call of the hot path over one batch: NightmareV3Env.step for E envs per GPU
(configs[1] of BASELINE.json: 4096 envs, flat ground, random actions), i.e. `decimation`=2 physics
substeps + the env epilogue, in one kernel launch.  Prints ONE JSON line (rank 0).

* value      : env-steps/s, whole job (all ranks), inputs resident in HBM, device-timed (CUDA events per
               step, max over ranks), L2 flushed between timed steps.
* e2e        : same metric through the public API (NightmareV3Env.step) with pinned HOST action buffers and
               a device->host read of obs/rew/done every step.
* roofline   : algorithmic HBM bytes per launch / event-timed kernel duration vs MEASURED_PEAKS.json; the
               kernel is FP32-pipe/latency bound, so `roofline_fp32` reports the FLOP view against an FFMA
               micro-benchmark measured in the same run.
* cpu_baseline: the fp64 CPU oracle (a restatement, "port"; real MuJoCo is not installable) on all host cores.
* --impl reference: times that CPU implementation as the reference arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")
METRIC = "hexapod env-steps/sec"
UNIT = "env-steps/s"
# Algorithmic work per env-step: fp32 words moved once per step and FLOPs of the
# restated pipeline (2 substeps).  The FLOP figure is the executed FP32 count per env-step measured with
# ncu on this kernel (profiles/), not an estimate of MuJoCo's own count.
BYTES_PER_ENV_STEP = 1464
FLOPS_PER_ENV_STEP = 71_800
# dram__bytes_read.sum + dram__bytes_write.sum of nm_step_kernel<true> per 4096-env launch (ncu --set full, profiles/r01_notes.md)
NCU_DRAM_BYTES_PER_LAUNCH_4096 = 11_502_848


def _workload(envs_per_gpu, decimation=2):
    return (f"nightmare_v3 {envs_per_gpu} envs/GPU, flat ground, N(0,1) random actions, decimation {decimation}, "
            f"random initial episode lengths (BASELINE configs[1])")


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_env_steps_per_s(n_envs, steps, warmup, threads, seed=1):
    """Times the CPU implementation of the path (oracle/: fp64 restatement of env.step + mj_step)."""
    import numpy as np
    from nightmare_rl_b200.envcfg import build_envcfg
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    from oracle import oracle as O
    cfg = NightmareV3Config()
    cfg.env.num_envs = n_envs
    om = O.OracleModel(NMB)
    b = O.OracleBatch(om, n_envs, seed=seed, envcfg=build_envcfg(cfg, 0.008))
    b.env_reset_idx(np.arange(n_envs))
    rng = np.random.default_rng(seed)
    b.env_set("ep_len", rng.integers(0, 1250, n_envs).astype(np.float64))
    acts = rng.normal(size=(8, n_envs, 18)).astype(np.float32)
    for i in range(warmup):
        b.env_step(acts[i % 8], threads)
    t0 = time.perf_counter()
    for i in range(steps):
        b.env_step(acts[i % 8], threads)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.envs_per_gpu
    steps = max(1, min(args.steps, 800))           # bounded (~10 s): each step is one pass over the whole 4096-env batch
    val, per = cpu_env_steps_per_s(n, steps, min(args.warmup, 2), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
        "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload(n), "envs": n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} envs x {steps} env-steps, fp64 C restatement of env.step+mj_step (MuJoCo itself is not installable), "
                                   f"{cores} pthreads over contiguous env slices"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the environment step has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from nightmare_rl_b200 import _lib
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env

    E, K, W = args.envs_per_gpu, args.steps, max(args.warmup, 3)
    cfg = NightmareV3Config()
    cfg.env.num_envs = E
    cfg.env.model_path = NMB
    cfg.viewer.render = False
    cfg.viewer.record_states = False
    env = NightmareV3Env(cfg, seed=1, env_offset=rank * E, device=dev)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    env.episode_length_buf = torch.randint(0, 1250, (E,), device=dev, generator=gen)      # init_at_random_ep_len (train.py:54)
    pool = torch.randn(16, E, 18, device=dev, generator=gen)                              # synthetic N(0,1) actions, resident in HBM
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                         # > 126 MB L2
    batch = env._batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_no = [env.common_step_counter]

    def one_step(i):
        step_no[0] += 1
        batch.step(pool[i % 16], step_no[0])

    for i in range(W):
        one_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = batch.launches
    barrier()
    t_wall0 = time.time()
    for i in range(K):
        flush.zero_()                              # evict the (L2-resident) state between timed steps; outside the event pair
        ev[i][0].record()
        one_step(W + i)
        ev[i][1].record()
    barrier()
    t_wall1 = time.time()
    launches = batch.launches - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = world * E * K / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers
    h_act = torch.randn(16, E, 18).pin_memory()
    h_obs, h_rew, h_done = torch.empty(E, 66).pin_memory(), torch.empty(E).pin_memory(), torch.empty(E, dtype=torch.int64).pin_memory()
    Ke = min(K, 500)
    for i in range(3):
        env.step_host(h_act[i % 16], h_obs, h_rew, h_done)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ke):
        # public host-side call: pinned actions H2D, step, obs/rew/dones D2H, stream sync -- every step
        obs, _, rew, done, _ = env.step_host(h_act[i % 16], h_obs, h_rew, h_done)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * E * Ke / (float(e2e_ms.item()) * 1e-3)

    # ---- batch-size sweep (metric is quoted on 4096-131072 envs): device-timed, state resident, rank 0 of a 1-GPU run only
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        for En in (16384, 65536, 131072):
            c2 = NightmareV3Config()
            c2.env.num_envs = En
            c2.env.model_path = NMB
            c2.viewer.render = False
            c2.viewer.record_states = False
            e2 = NightmareV3Env(c2, seed=1, device=dev)
            e2.reset()
            e2.episode_length_buf = torch.randint(0, 1250, (En,), device=dev, generator=gen)
            acts = torch.randn(4, En, 18, device=dev, generator=gen)
            for i in range(5):
                e2._batch.step(acts[i % 4], 100 + i)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            Ks = 30
            s0.record()
            for i in range(Ks):
                e2._batch.step(acts[i % 4], 200 + i)
            s1.record()
            torch.cuda.synchronize()
            ms = s0.elapsed_time(s1) / Ks
            sweep.append({"envs": En, "ms_per_step": ms, "env_steps_per_s": En / (ms * 1e-3)})
            del e2, acts
            torch.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, which = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        kern_ms = sum(step_ms) / K                                       # one launch per step: event pair brackets exactly that launch
        gbs = BYTES_PER_ENV_STEP * E / (kern_ms * 1e-3) / 1e9
        fp32_peak = _lib.lib.nm_measure_fp32_peak(None)
        tfs = FLOPS_PER_ENV_STEP * E / (kern_ms * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        cpu_steps = 800                                   # ~10 s of CPU work on 16 cores
        cpu_val, _ = cpu_env_steps_per_s(E, cpu_steps, 2, cores)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": _workload(E), "envs_per_gpu": E, "envs_total": world * E, "parallelism": f"env-sharded x{world}",
                       "l2": "flushed between timed steps (256 MiB memset outside the event pair)"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": E * 18 * 4, "d2h_bytes_per_step": E * (66 * 4 + 4 + 8), "steps": Ke},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH_4096 if E == 4096 else None,
                         "algorithmic_bytes": BYTES_PER_ENV_STEP * E,
                         "peak_source": which, "kernel": "nm_step_kernel<true> (+ the 2 us nm_finalize_kernel inside the same event pair)", "kernel_ms": kern_ms,
                         "note": "kernel is FP32-pipe/latency bound, not HBM bound; see roofline_fp32"},
            "roofline_fp32": {"bound": "fp32", "achieved": tfs, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tfs / fp32_peak if fp32_peak > 0 else None,
                              "flops_per_env_step": FLOPS_PER_ENV_STEP, "peak_source": "FFMA micro-benchmark in this run"},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{E} envs x {cpu_steps} env-steps, fp64 C restatement (oracle/), {cores} pthreads"},
        }
        if sweep is not None:
            line["sweep"] = {"note": "same step at larger batches on 1 GPU, back-to-back launches, no L2 flush", "points": sweep}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--no-sweep", action="store_true", help="skip the 16384/65536/131072-env sweep of the 1-GPU run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
