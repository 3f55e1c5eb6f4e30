#!/usr/bin/env python
"""bench.py — hexapod env-steps/s of the batched Nightmare-v3 environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload step|rollout|ppo|anymal_c] [--envs-per-gpu E]

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
* rollout / ppo: sub-records of the same line for BASELINE configs[2] (16 384 envs/GPU, domain randomisation, PPO rollout with
               the tcgen05 policy forward) and configs[4] (16 384 envs/GPU = 131 072 over 8 GPUs, full PPO iteration); with
               N > 1 ranks the ppo record times the in-graph NCCL all-reduces (update with vs without them).
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
# Algorithmic work per env-step (DESIGN.md §4 "Roofline").
#  * bytes: fp32 words that have to move once per step with the whole step fused in one launch.
#  * flops: counted by the INSTRUMENTED ORACLE (oracle/libnm_oracle_cnt.so: every add/mul/div/sqrt of the restated pipeline,
#    structural zeros skipped the way a tree-sparse implementation skips them) on a CPU sample of this very workload, in this
#    run -- not the kernel's own executed count, which is kept as a second field (`executed_flops_per_env_step`, from the
#    newest ncu capture under profiles/).
BYTES_PER_ENV_STEP = 1464
PREROLL_STEPS = 300       # untimed env steps before the warm-up: robots have landed and are in the steady random-action regime
FLOP_STAGES = ("kinematics", "comPos", "crb", "factorM", "collision", "makeConstraint", "projectConstraint", "comVel+rne", "actuation+qacc_smooth",
               "warmstart", "PGS", "noslip", "qfrc_constraint+qacc", "touch sensors", "implicitfast integrate", "env layer")


def _profile_numbers():
    """dram bytes per launch and executed FP32 flops per env-step of nm_step_kernel<true> from the newest ncu summary in
    profiles/ (written by tools/update_profile.py from a `ncu --set full` capture of `python bench.py --steps 2 --warmup 1`)."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_kernel_metrics.txt"))):
        best = path
    if best is None:
        return None
    txt = open(best).read()

    def num(key):
        m = re.search(re.escape(key) + r"\s*=\s*([0-9.eE+-]+)\s*(\w*)", txt)
        if not m:
            return None
        v = float(m.group(1))
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(m.group(2), 1.0)

    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    grid, block = num("launch__grid_size"), num("launch__block_size")
    ffma = num("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed")
    fadd = num("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed")
    fmul = num("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")
    cyc = num("sm__cycles_elapsed.max")
    out = {"file": os.path.relpath(best, ROOT), "envs": int(grid * block / 8) if grid and block else None,
           "dram_bytes": (rd + wr) if rd is not None and wr is not None else None, "executed_flops_per_env_step": None}
    out["grid"], out["envs_per_cta"] = (int(grid), int(block) // 8) if grid and block else (None, None)
    if None not in (ffma, fadd, fmul, cyc) and out["envs"]:
        out["executed_flops_total"] = (2 * ffma + fadd + fmul) * cyc
        out["executed_flops_per_env_step"] = out["executed_flops_total"] / out["envs"]
    return out


def _profile_matches(prof, E):
    """the captured launch is this batch size: the same number of CTAs (the last CTA may be partly filled)"""
    return bool(prof.get("grid")) and prof["grid"] == -(-E // prof["envs_per_cta"])


def _config(E, world, workload=None):
    """`config` of the JSON line -- one function for both arms, so the driver sees identical dicts."""
    return {"workload": workload or _workload(E), "envs_per_gpu": E, "envs_total": world * E, "parallelism": f"env-sharded x{world}",
            "l2": "flushed between timed steps (256 MiB memset outside the event pair)",
            "preroll": f"{PREROLL_STEPS} untimed env steps before the warm-up (robots landed, steady random-action regime)"}


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
def _cpu_batch(n_envs, seed, variant="f64", preroll=PREROLL_STEPS, threads=1):
    """Oracle batch brought into the workload's regime: reset, random episode lengths, `preroll` random-action steps."""
    import numpy as np
    from nightmare_rl_b200.envcfg import build_envcfg
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    from oracle import oracle as O
    cfg = NightmareV3Config()
    cfg.env.num_envs = n_envs
    om = O.OracleModel(NMB, variant=variant)
    b = O.OracleBatch(om, n_envs, seed=seed, envcfg=build_envcfg(cfg, 0.008))
    b.env_reset_idx(np.arange(n_envs))
    rng = np.random.default_rng(seed)
    b.env_set("ep_len", rng.integers(0, 1250, n_envs).astype(np.float64))
    acts = rng.normal(size=(8, n_envs, 18)).astype(np.float32)
    for i in range(preroll):
        b.env_step(acts[i % 8], threads)
    return om, b, acts


def cpu_env_steps_per_s(n_envs, steps, warmup, threads, seed=1, reps=1):
    """Times the CPU implementation of the path (oracle/: fp64 restatement of env.step + mj_step), all host threads over
    contiguous env slices the way simple_test.py:25-45 threads MuJoCo.  One timed "step" = `reps` consecutive env steps."""
    _, b, acts = _cpu_batch(n_envs, seed, threads=threads)
    for i in range(warmup * reps):
        b.env_step(acts[i % 8], threads)
    t0 = time.perf_counter()
    for i in range(steps * reps):
        b.env_step(acts[i % 8], threads)
    dt = time.perf_counter() - t0
    return n_envs * steps * reps / dt, dt / steps


def oracle_flops_per_env_step(n_envs=192, steps=16, seed=1):
    """Algorithmic FLOPs of one env step: the instrumented oracle (oracle/libnm_oracle_cnt.so) counts every floating-point
    operation of the restated pipeline on a sample of this workload (same reset / pre-roll / action distribution)."""
    import ctypes
    import numpy as np
    om, b, acts = _cpu_batch(n_envs, seed, variant="cnt", threads=1)
    om.L.nmo_flop_reset()
    for i in range(steps):
        b.env_step(acts[i % 8], 1)
    out = np.zeros(len(FLOP_STAGES))
    om.L.nmo_flop_counts(out.ctypes.data_as(ctypes.c_void_p), out.size)
    per = out / (n_envs * steps)
    ncon = float(np.mean([b.get(i, "ncon")[0] for i in range(n_envs)]))
    return float(per.sum()), {k: round(float(v)) for k, v in zip(FLOP_STAGES, per)}, ncon


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    cores = os.cpu_count() or 1
    n = args.envs_per_gpu
    K, W = max(1, args.steps), max(args.warmup, 0)
    # one reference-arm "step" is a bounded sample: `reps` passes of the n-env batch, so that short --steps runs still time
    # >= 200 env steps (the CPU arm needs ~1 s to reach its steady rate) and long ones stay within a few minutes
    reps = max(1, -(-200 // K)) if K < 200 else 1
    if K * reps > 2000:
        reps = 1
    val, per = cpu_env_steps_per_s(n, K, W, cores, reps=reps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(n, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} envs x {K} steps of {reps} env-step(s) each (+ {W} warm-up steps, {PREROLL_STEPS} pre-roll env steps), fp64 C "
                                   f"restatement of env.step+mj_step (MuJoCo itself is not installable), {cores} pthreads over contiguous env slices"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
class _Ctx:
    pass


def _make_env(E, rank, dev, seed=1, dr=False):
    import torch
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env
    cfg = NightmareV3Config()
    cfg.env.num_envs = E
    cfg.env.model_path = NMB
    cfg.viewer.render = False
    cfg.viewer.record_states = False
    env = NightmareV3Env(cfg, seed=seed, env_offset=rank * E, device=dev)
    if dr:
        env.set_domain_randomization(friction=(0.5, 1.25), kv=(0.8, 1.2), base_mass=(-0.3, 0.3), resample_on_reset=True)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    env.episode_length_buf = torch.randint(0, 1250, (E,), device=dev, generator=gen)      # init_at_random_ep_len (train.py:54)
    return env, gen


def bench_step(c, args):
    """BASELINE configs[1]: the env step alone, random actions resident in HBM."""
    import torch
    import torch.distributed as dist
    dev, world, rank = c.dev, c.world, c.rank
    E, K, W = args.envs_per_gpu, args.steps, max(args.warmup, 3)
    env, gen = _make_env(E, rank, dev)
    pool = torch.randn(16, E, 18, device=dev, generator=gen)                              # synthetic N(0,1) actions, resident in HBM
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                         # > 126 MB L2
    batch = env._batch
    step_no = [env.common_step_counter]

    def one_step(i):
        step_no[0] += 1
        batch.step(pool[i % 16], step_no[0])

    for i in range(PREROLL_STEPS):
        one_step(i)
    c.barrier()
    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)                           # the sampler thread is up before anything is timed ...
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    c.barrier()
    for i in range(W):                             # ... and the W warm-up steps run right before the timed ones, with the same
        flush.zero_()                              # flush in between (a first step after an idle quarter second costs 1.6x)
        one_step(i)
    l0 = batch.launches
    c.barrier()
    t_wall0 = time.time()
    for i in range(K):
        flush.zero_()                              # evict the (L2-resident) state between timed steps; outside the event pair
        ev[i][0].record()
        one_step(W + i)
        ev[i][1].record()
    c.barrier()
    t_wall1 = time.time()
    launches = batch.launches - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = c.max_over_ranks(sum(step_ms))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = world * E * K / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers
    h_act = torch.randn(16, E, 18).pin_memory()
    h_obs, h_rew, h_done = torch.empty(E, 66).pin_memory(), torch.empty(E).pin_memory(), torch.empty(E, dtype=torch.int64).pin_memory()
    Ke = max(K, min(200, 10 * K))                 # short driver runs (--steps 20) still time a few hundred public-API calls
    for i in range(3):
        env.step_host(h_act[i % 16], h_obs, h_rew, h_done)
    c.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ke):
        # public host-side call: pinned actions H2D, step, obs/rew/dones D2H, stream sync -- every step
        env.step_host(h_act[i % 16], h_obs, h_rew, h_done)
    e1.record()
    c.barrier()
    e2e_ms = c.max_over_ranks(e0.elapsed_time(e1))
    e2e_val = world * E * Ke / (e2e_ms * 1e-3)

    # ---- batch-size sweep (metric is quoted on 4096-131072 envs): device-timed, state resident, 1-GPU run only
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        for En in (16384, 65536, 131072):
            e2, g2 = _make_env(En, 0, dev)
            acts = torch.randn(4, En, 18, device=dev, generator=g2)
            for i in range(PREROLL_STEPS + 5):
                e2._batch.step(acts[i % 4], 100 + i)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            Ks = 30
            s0.record()
            for i in range(Ks):
                e2._batch.step(acts[i % 4], 1000 + i)
            s1.record()
            torch.cuda.synchronize()
            ms = s0.elapsed_time(s1) / Ks
            sweep.append({"envs": En, "ms_per_step": ms, "env_steps_per_s": En / (ms * 1e-3)})
            del e2, acts
            torch.cuda.empty_cache()
    # ---- the same step on a walking population (the reference's scripted tripod gait, phase-shifted per env): what a trained
    # policy's rollouts look like -- six feet in contact, adjacent tibias close enough to enter the pair broad phase
    walking = None
    if world == 1 and not args.no_sweep:
        from nightmare_rl_b200.envs.scripted_gait import ScriptedGait
        e3, _ = _make_env(E, 0, dev)
        gait = ScriptedGait(os.path.join(ROOT, "tests", "golden", "nikengine_gait_targets.npz"), E, dev, phase_shift=3)
        settle, Kw = 330, 50
        seq = [gait.actions().contiguous() for _ in range(settle + Kw)]
        for i in range(settle):
            e3._batch.step(seq[i], 100 + i)
        torch.cuda.synchronize()
        evw = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kw)]
        for i in range(Kw):
            flush.zero_()
            evw[i][0].record()
            e3._batch.step(seq[settle + i], 1000 + i)
            evw[i][1].record()
        torch.cuda.synchronize()
        msw = sum(a.elapsed_time(b) for a, b in evw) / Kw
        walking = {"workload": "same batch, actions from the reference's scripted tripod gait (tests/golden/nikengine_gait_targets.npz), "
                               "330 untimed steps, L2 flushed between timed steps", "envs": E, "steps": Kw, "ms_per_step": msw,
                   "env_steps_per_s": E / (msw * 1e-3)}
        del e3, seq
    # ---- the other way the timing rules allow to keep inputs out of L2: a working set larger than L2.  NB batches of E envs are
    # stepped round-robin (their state, carried buffers and outputs: > 200 MB, L2 is 126 MB), so every step reads its inputs from
    # HBM -- while the kernel's 256 KB of instructions stay L2-resident between launches, as they do in a training loop.  The
    # headline `value` flushes L2 with a 256 MiB memset instead, which also evicts the code (every instruction line of a step then
    # comes from DRAM once); the difference between the two records is that refetch.
    rotating = None
    if world == 1 and not args.no_sweep:
        per_batch = E * (25 + 24 + 24 + 18 * 4 + 3 + 19 + 6 + 66 + 1 + 2 + 2 + 13 + 8) * 4      # bytes of state + carried buffers + outputs
        NB = max(8, -(-(200 << 20) // per_batch))
        envs_r = [_make_env(E, 0, dev, seed=100 + k)[0] for k in range(NB)]
        acts_r = torch.randn(4, E, 18, device=dev, generator=gen)
        for i in range(PREROLL_STEPS):
            for er in envs_r:
                er._batch.step(acts_r[i % 4], 10 + i)
        torch.cuda.synchronize()
        Kr = 4 * NB
        evr = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kr)]
        for i in range(Kr):
            evr[i][0].record()
            envs_r[i % NB]._batch.step(acts_r[i % 4], 1000 + i)
            evr[i][1].record()
        torch.cuda.synchronize()
        msr = sorted(a.elapsed_time(b) for a, b in evr)
        rotating = {"workload": f"{NB} batches of {E} envs stepped round-robin ({NB * per_batch / 2**20:.0f} MiB of state, buffers and outputs > 126 MB L2), "
                                "no flush: inputs from HBM, instructions L2-resident", "batches": NB, "steps": Kr,
                    "ms_per_step": sum(msr) / Kr, "ms_median": msr[Kr // 2], "env_steps_per_s": E * Kr / (sum(msr) * 1e-3)}
        del envs_r, acts_r
    del flush
    torch.cuda.empty_cache()
    srt = sorted(step_ms)
    return dict(walking=walking, rotating=rotating, value=value, total_ms=total_ms, kern_ms=sum(step_ms) / K, clocks=clocks, launches=int(launches), e2e_val=e2e_val, Ke=Ke,
                sweep=sweep, E=E, K=K, W=W, step_ms_stats={"min": srt[0], "median": srt[len(srt) // 2], "max": srt[-1]})


def bench_rollout(c, E=16384, T=80, reps=3):
    """BASELINE configs[2]: E envs per GPU with domain randomisation + the PPO rollout (tcgen05 policy forward, env step,
    transition store), 80 steps per rollout (envs/nightmare_v3_config.py:135)."""
    import tempfile
    import torch
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3ConfigPPO
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    dev = c.dev
    tc = NightmareV3ConfigPPO()
    env, _ = _make_env(E, c.rank, dev, seed=tc.seed, dr=True)
    torch.manual_seed(tc.seed)
    ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
    alg = PPO(ac, gamma=0.99, device=str(dev), fused_rollout=True, graph_update=False, seed=1, env_offset=c.rank * E)
    alg.init_storage(E, T, [66], [None], [18])
    alg.attach_episode_stats(torch.zeros(E, device=dev), torch.zeros(E, device=dev), torch.zeros(100, device=dev), torch.zeros(100, device=dev),
                             torch.zeros(1, dtype=torch.int64, device=dev))
    if not alg.prepare_fast_rollout(env, torch.zeros(32, device=dev)):
        raise RuntimeError("fused rollout path unavailable")
    for _ in range(PREROLL_STEPS // T + 1):                  # pre-roll under the (random-init) policy
        alg.storage.clear()
        for t in range(T):
            alg.fast_rollout_step()
    times = []
    l0 = env.gpu_launches
    for rep in range(reps):
        alg.storage.clear()
        c.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(T):
            alg.fast_rollout_step()
        e1.record()
        c.barrier()
        times.append(c.max_over_ranks(e0.elapsed_time(e1)))
    ms = sorted(times)[len(times) // 2]
    rec = {"workload": f"nightmare_v3 {E} envs/GPU, domain randomisation (friction U(0.5,1.25), kv x U(0.8,1.2), base mass +-0.3 kg, redrawn on reset), "
                       f"{T}-step PPO rollout: tcgen05 policy forward + env step + transition store (BASELINE configs[2])",
           "envs_per_gpu": E, "envs_total": c.world * E, "rollout_steps": T, "value": c.world * E * T / (ms * 1e-3), "unit": UNIT,
           "ms_per_rollout": ms, "us_per_step": ms / T * 1e3, "policy_engine": alg.fused.engine, "reps": reps,
           "gpu_launches_per_step": (env.gpu_launches - l0) / (reps * T) + 2}
    del alg, env
    torch.cuda.empty_cache()
    return rec


def bench_ppo(c, E=16384, iters=3):
    """BASELINE configs[4]: E envs per GPU (131 072 over 8 GPUs), the full train.py loop -- 80-step rollout, GAE, 5 epochs x 4
    mini-batches (envs/nightmare_v3_config.py:123-135), KL-adaptive Adam -- with the gradient / KL / advantage-moment
    all-reduces over NCCL when world > 1.  The collective's cost is measured by running the same update with and without it."""
    import tempfile
    import torch
    import torch.distributed as dist
    from envs.helpers import class_to_dict
    from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
    from envs.nightmare_v3_env import NightmareV3Env
    from rsl_rl.runners import OnPolicyRunner
    dev = c.dev
    cfg, tc = NightmareV3Config(), NightmareV3ConfigPPO()
    cfg.env.num_envs = E
    cfg.env.model_path = NMB
    cfg.viewer.render = cfg.viewer.record_states = False
    env = NightmareV3Env(cfg, log_dir=tempfile.mkdtemp(), seed=tc.seed, env_offset=c.rank * E, device=dev)
    runner = OnPolicyRunner(env, class_to_dict(tc), log_dir=None, device=str(dev))
    T = runner.num_steps_per_env
    runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)       # warm-up: graph capture, allocator, NCCL

    def timed(n):
        coll, learn = 0.0, 0.0
        c.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            runner.learn(num_learning_iterations=1)
            coll += runner.last_log["collection_time"]
            learn += runner.last_log["learn_time"]
        e1.record()
        c.barrier()
        return c.max_over_ranks(e0.elapsed_time(e1)) / n, c.max_over_ranks(coll * 1e3) / n, c.max_over_ranks(learn * 1e3) / n

    it_ms, roll_ms, upd_ms = timed(iters)
    rec = {"workload": f"nightmare_v3 {E} envs/GPU, full PPO iteration (train.py:54): {T}-step rollout + GAE + "
                       f"{runner.alg.num_learning_epochs} epochs x {runner.alg.num_mini_batches} mini-batches, KL-adaptive Adam"
                       + (f", NCCL all-reduce of gradient+KL per mini-batch over {c.world} GPUs (BASELINE configs[4])" if c.world > 1 else " (BASELINE configs[4] at 1 GPU)"),
           "envs_per_gpu": E, "envs_total": c.world * E, "value": c.world * E * T / (it_ms * 1e-3), "unit": UNIT, "ms_per_iteration": it_ms,
           "rollout_ms": roll_ms, "update_ms": upd_ms, "rollout_env_steps_per_s": c.world * E * T / (roll_ms * 1e-3), "iterations": iters,
           "update_engine": "fused nm_ppo_grad + nm_ppo_adam (CUDA graph)" if runner.alg.fused_grad is not None else "autograd (CUDA graph)"}
    if c.world > 1:
        # the same update without the collectives (every rank on its own shard; parameters diverge from here on, nothing is
        # measured after this): the difference is what the in-graph all-reduces cost per iteration, skew between ranks included
        n_coll = runner.alg.num_learning_epochs * runner.alg.num_mini_batches
        runner.alg.release_graph()
        runner.alg.world = 1
        runner.learn(num_learning_iterations=1)
        _, _, upd_nc = timed(iters)
        # and the collective alone: n_coll back-to-back all-reduces of the same flat buffer
        buf = torch.zeros(runner.alg.fused_grad.ext.numel() if runner.alg.fused_grad is not None else 15043 + 8, device=dev)
        for _ in range(5):
            dist.all_reduce(buf)
        c.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_coll):
            dist.all_reduce(buf)
        e1.record()
        c.barrier()
        alone = c.max_over_ranks(e0.elapsed_time(e1))
        rec["allreduce"] = {"collectives_per_iteration": n_coll + 1, "bytes_each": int(buf.numel() * 4), "update_ms_with": upd_ms, "update_ms_without": upd_nc,
                            "in_graph_cost_ms_per_iteration": upd_ms - upd_nc, "back_to_back_ms_per_iteration": alone,
                            "us_per_allreduce_back_to_back": alone / n_coll * 1e3, "backend": "nccl"}
    runner.alg.release_graph()
    del runner, env
    torch.cuda.empty_cache()
    return rec


def bench_anymal(c, E=4096, K=200):
    """BASELINE configs[3]: the reference's second model (models/anymal_c: Newton solver, elliptic cones, condim-6 feet,
    friction loss, joint limits, position actuators) -- E envs per GPU, dt 0.002, 4 substeps per env step
    (≙ mj.mj_step(model, data[i], 4), simple_test.py:39), joint targets redrawn every step from U(-0.35, 0.35) rad around the
    standing pose; robots whose trunk drops below 0.3 m or tilts past 60 deg are put back on their feet (torch ops inside the
    timed region), so the population stays in the standing / stumbling regime."""
    import torch
    from nightmare_rl_b200 import _lib, mjcf
    from nightmare_rl_b200.batch import GenBatch
    dev = c.dev
    path = os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb")
    cm = mjcf.CompiledModel.load(path)
    gb = GenBatch(_lib.GenModel(cm.to_bytes()), E, dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + c.rank)
    q0 = gb.qpos[0].clone()
    pool = (torch.rand(16, E, 12, device=dev, generator=gen) - 0.5) * 0.7
    resets = torch.zeros(1, dtype=torch.int64, device=dev)

    def one(i):
        gb.physics_step(pool[i % 16], 4)
        qw, qz = gb.qpos[:, 3], gb.qpos[:, 6]
        fallen = (gb.qpos[:, 2] < 0.3) | (1.0 - 2.0 * (gb.qpos[:, 4] ** 2 + gb.qpos[:, 5] ** 2) < 0.5)      # trunk z axis . world z < cos 60
        gb.qpos[fallen] = q0
        gb.qvel[fallen] = 0.0
        resets.add_(fallen.sum())

    for i in range(300):
        one(i)
    resets.zero_()
    ncon_mean = float(gb.info[:, 0].float().mean())
    nefc_mean, it_mean, ovf = float(gb.info[:, 1].float().mean()), float(gb.info[:, 2].float().mean()), int(gb.info[:, 3].sum())
    c.barrier()
    l0 = gb.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        one(i)
    e1.record()
    c.barrier()
    ms = c.max_over_ranks(e0.elapsed_time(e1)) / K
    # the physics launches alone (no reset bookkeeping), same state distribution
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(50):
        gb.physics_step(pool[i % 16], 4)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / 50
    rec = {"workload": f"anymal_c {E} envs/GPU, dt 0.002 x 4 substeps per env step, random joint targets U(-0.35, 0.35) rad, fallen robots re-seated "
                       "(BASELINE configs[3]; Newton solver, elliptic cones impratio 100, condim-6 feet, friction loss, joint limits)",
           "envs_per_gpu": E, "envs_total": c.world * E, "value": c.world * E / (ms * 1e-3), "unit": UNIT, "substeps_per_s": 4 * c.world * E / (ms * 1e-3),
           "ms_per_step": ms, "physics_launch_ms": kern_ms, "steps": K, "gpu_launches_per_step": (gb.launches - l0) / K,
           "resets_per_step": float(resets.item()) / K, "mean_contacts": ncon_mean, "mean_constraint_rows": nefc_mean, "mean_newton_iterations_last_substep": it_mean,
           "truncated_envs": ovf, "dtype": "f32 (+ fp64 Newton iterate / residual)",
           "algorithmic_bytes_per_env_step": 4 * (19 + 18 + 18 + 12 + 19 + 18 + 18 + 4),
           "hbm_gbs": 4 * (19 + 18 + 18 + 12 + 19 + 18 + 18 + 4) * E / (kern_ms * 1e-3) / 1e9}
    if c.rank == 0 and c.world == 1:
        # CPU arm on the same workload: the oracle (fp64 restatement of mj_step for this model) on all host cores, bounded sample
        import numpy as np
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        n_cpu = min(E, 512)
        ob = O.OracleBatch(O.OracleModel(path), n_cpu)
        rng = np.random.default_rng(7)
        ctrl = (rng.random((n_cpu, 12)) - 0.5) * 0.7
        ob.physics_step(ctrl, 100, cores)
        t0 = time.time()
        nst = 10
        for _ in range(nst):
            ob.physics_step(ctrl, 4, cores)
        dt = time.time() - t0
        rec["cpu_baseline"] = {"value": n_cpu * nst / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{n_cpu} envs x {nst} env steps of 4 substeps after 100 settle substeps, fp64 oracle (exact Newton), {cores} pthreads"}
    del gb
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    c = _Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", 1))
    c.rank = int(os.environ.get("RANK", 0))
    c.local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the environment step has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=c.dev, dtype=torch.float64)
        if c.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c.barrier, c.max_over_ranks = barrier, max_over_ranks
    from nightmare_rl_b200 import _lib

    world, rank = c.world, c.rank
    st = bench_step(c, args)
    E, K, W = st["E"], st["K"], st["W"]
    extras = {}
    if not args.no_extras:
        # BASELINE configs[2] and configs[4] in the same line, so that the driver's runs (1 GPU and the 1/2/4/8 scaling run)
        # carry a rollout number and a full-PPO number whose timed region contains the NCCL collectives
        for name, fn in (("rollout", bench_rollout), ("ppo", bench_ppo), ("anymal_c", bench_anymal)):
            try:
                extras[name] = fn(c, E=args.envs_per_gpu if name == "anymal_c" else args.extras_envs_per_gpu)
            except Exception as exc:              # the headline must survive a failing extra; the failure is reported, not hidden
                extras[name] = {"error": f"{type(exc).__name__}: {exc}"}
                barrier()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, which = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        kern_ms = st["kern_ms"]                                          # one launch per step: event pair brackets exactly that launch
        gbs = BYTES_PER_ENV_STEP * E / (kern_ms * 1e-3) / 1e9
        fp32_peak = _lib.lib.nm_measure_fp32_peak(None)
        flops, stages, ncon = oracle_flops_per_env_step()
        prof = _profile_numbers() or {}
        tfs = flops * E / (kern_ms * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        cpu_steps = 600                                   # ~10 s of CPU work on 16 cores
        cpu_val, _ = cpu_env_steps_per_s(E, cpu_steps, 2, cores)
        traffic = prof.get("dram_bytes") if _profile_matches(prof, E) else None
        if _profile_matches(prof, E) and prof.get("executed_flops_total"):
            prof["executed_flops_per_env_step"] = prof["executed_flops_total"] / E
        line = {
            "metric": METRIC, "value": st["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": st["total_ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": _config(E, world),
            "clocks": st["clocks"],
            "e2e": {"value": st["e2e_val"], "unit": UNIT, "h2d_bytes_per_step": E * 18 * 4, "d2h_bytes_per_step": E * (66 * 4 + 4 + 8), "steps": st["Ke"]},
            "gpu_launches": st["launches"],
            "step_ms_stats": st["step_ms_stats"],
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
                         "traffic_source": prof.get("file"), "algorithmic_bytes": BYTES_PER_ENV_STEP * E,
                         "peak_source": which, "kernel": "nm_step_kernel<true> (+ the 2 us nm_finalize_kernel inside the same event pair)", "kernel_ms": kern_ms,
                         "note": "kernel is FP32-pipe/latency bound, not HBM bound; see roofline_fp32"},
            "roofline_fp32": {"bound": "fp32", "achieved": tfs, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tfs / fp32_peak if fp32_peak > 0 else None,
                              "flops_per_env_step": flops, "flops_source": "instrumented oracle (oracle/libnm_oracle_cnt.so) on a CPU sample of this workload, this run",
                              "flops_by_stage": stages, "mean_contacts_per_env": ncon,
                              "executed_flops_per_env_step": prof.get("executed_flops_per_env_step"), "executed_source": prof.get("file"),
                              "peak_source": "FFMA micro-benchmark in this run"},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{E} envs x {cpu_steps} env-steps after {PREROLL_STEPS} pre-roll steps, fp64 C restatement (oracle/), {cores} pthreads"},
        }
        if st.get("walking") is not None:
            line["walking"] = st["walking"]
        if st.get("rotating") is not None:
            line["rotating_inputs"] = st["rotating"]
        if st["sweep"] is not None:
            line["sweep"] = {"note": "same step at larger batches on 1 GPU, back-to-back launches, no L2 flush", "points": st["sweep"]}
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_workload(args):
    """`--workload rollout|ppo`: that workload alone as the headline of the line."""
    import torch
    import torch.distributed as dist
    c = _Ctx()
    c.world, c.rank, c.local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=c.dev, dtype=torch.float64)
        if c.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c.barrier, c.max_over_ranks = barrier, max_over_ranks
    E = args.extras_envs_per_gpu
    if args.workload == "anymal_c":
        E = args.envs_per_gpu
        rec = bench_anymal(c, E=E, K=max(20, min(args.steps, 1000)))
    else:
        rec = bench_rollout(c, E=E, reps=max(3, min(args.steps, 10))) if args.workload == "rollout" else bench_ppo(c, E=E, iters=max(3, min(args.steps, 10)))
    if c.rank == 0:
        ms = rec.get("ms_per_rollout", rec.get("ms_per_iteration", rec.get("ms_per_step")))
        line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": c.world, "steps": rec.get("reps", rec.get("iterations")), "warmup": 2,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": _config(E, c.world, rec["workload"]), args.workload: rec}
        print(json.dumps(line), flush=True)
    if c.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["step", "rollout", "ppo", "anymal_c"], default="step",
                    help="step (default, the headline: BASELINE configs[1]; its line also carries `rollout` and `ppo` sub-records), "
                         "rollout (configs[2]), ppo (configs[4]) or anymal_c (configs[3]) alone")
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--extras-envs-per-gpu", type=int, default=16384, help="envs per GPU of the rollout / ppo records (131072 over 8 GPUs)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 16384/65536/131072-env sweep of the 1-GPU run")
    ap.add_argument("--no-extras", action="store_true", help="skip the rollout / ppo sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "step":
        run_workload(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
