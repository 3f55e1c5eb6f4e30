#!/usr/bin/env python
"""Per-phase critical-path breakdown of nm_step_kernel from clock64() stamps (needs the -DNM_TIMING build exp/timing.so).

    NIGHTMARE_B200_LIB=$PWD/exp/timing.so python tools/phase_timing.py [N]
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nightmare_rl_b200 import _lib
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
cfg = NightmareV3Config(); cfg.env.num_envs = N; cfg.viewer.render = cfg.viewer.record_states = False
env = NightmareV3Env(cfg, seed=1, device=dev); env.reset()
env.episode_length_buf = torch.randint(0, 1250, (N,), device=dev)
acts = torch.randn(8, N, 18, device=dev)
GAIT = os.environ.get("NM_GAIT", "0") == "1"
if GAIT:                                                     # the reference's scripted tripod gait instead of N(0,1) actions
    from nightmare_rl_b200.envs.scripted_gait import ScriptedGait
    g = ScriptedGait(os.path.join(ROOT, "tests", "golden", "nikengine_gait_targets.npz"), N, dev, phase_shift=3)
    seq = [g.actions().contiguous() for _ in range(400)]
    for i in range(330):
        env._batch.step(seq[i], 10 + i)
    acts = seq[330:]                                         # the gait goes on through the sampled steps
for i in range(40):
    env._batch.step(acts[i % 8 if not GAIT else i], 10 + i)
nw = (N * 8 + 31) // 32
buf = torch.zeros(nw + 8, 32, dtype=torch.int64, device=dev)
_lib.lib.nm_debug_set_timing_buffer.argtypes = [ctypes.c_void_p]
assert _lib.lib.nm_debug_set_timing_buffer(buf.data_ptr()) == 0
names = {1: "prev tail", 2: "wait@top", 3: "P1+P2+P7 kin/RNE", 4: "wait+P3 CRBA", 5: "wait+P8 factor x2 + solve", 6: "wait+P4 collision",
         7: "wait+P5 contact build", 8: "P9 sweeps + P10", 9: "wait@end of contacts"}
acc = {}
for rep in range(10):
    buf.zero_()
    env._batch.step(acts[rep % 8 if not GAIT else 40 + rep], 100 + rep)
    torch.cuda.synchronize()
    t = buf[:nw].cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    def add(k, v): acc.setdefault(k, []).append(v)
    add("total (first start -> last end)", t[:, 22].max() - t0)
    add("launch skew (last warp start)", t[:, 0].max() - t0)
    add("prologue: const copy + state load", np.median(t[:, 1] - t[:, 0]))
    for sub in range(2):
        b = 10 * sub
        prev = t[:, 1 + b]
        for k in range(2, 10):
            cur = t[:, k + b]
            if k == 7:
                cur = np.where(cur > 0, cur, t[:, 6 + b])      # warps without contacts skip the stamp
            d = cur - prev
            add(f"sub{sub} {names[k]:28s}", (np.median(d), d.max()))
            prev = cur
    add("P11 of last substep + epilogue", np.median(t[:, 22] - t[:, 19]))
    if t[:, 23].max() > 0:                                   # finer stamps inside the collision phase of substep 0 (NM_TIMING builds)
        add("sub0 barrier before P4", (np.median(t[:, 23] - t[:, 5]), (t[:, 23] - t[:, 5]).max()))
        add("sub0 P4 hull-plane (walk + extra contacts)", (np.median(t[:, 25] - t[:, 23]), (t[:, 25] - t[:, 23]).max()))
        add("sub0   of which the support-vertex walk", (np.median(t[:, 24] - t[:, 23]), (t[:, 24] - t[:, 23]).max()))
        add("sub0   walk rounds per launch (max over the warp's hulls: median, max)", (np.median(t[:, 27]), t[:, 27].max()))
        if rep == 0:
            wt = (t[:, 24] - t[:, 23]) / 1965.0
            for r_ in sorted(set(t[:, 27].astype(int).tolist())):
                m_ = t[:, 27].astype(int) == r_
                print(f"  walk rounds {r_:2d}: warps {m_.sum():4d}  walk median {np.median(wt[m_]):5.2f} us  max {wt[m_].max():5.2f} us")
        add("sub0   max vertex degree at the support vertex", (np.median(t[:, 28]), t[:, 28].max()))
        add("sub0 P4b tibia pairs broad phase", (np.median(t[:, 26] - t[:, 25]), (t[:, 26] - t[:, 25]).max()))
        pk = t[:, 29].astype(np.int64)
        add("sub0   pair candidates per launch: capsules overlap / survive the support-map axis test", ((pk & 0xfffff).sum(), ((pk >> 20) & 0xfffff).sum()))
        add("sub0   pairs per launch handed to MPR (sum, max per warp)", ((pk >> 40).sum(), (pk >> 40).max()))
        add("sub0   candidate filter (support maps)", (np.median(t[:, 30][t[:, 30] > 0]) if (t[:, 30] > 0).any() else 0, t[:, 30].max()))
        add("sub0   MPR + contact blocks", (np.median(t[:, 31][t[:, 31] > 0]) if (t[:, 31] > 0).any() else 0, t[:, 31].max()))
        if os.environ.get("NM_VISIT") == "1":
            for nm_, col in (("PGS visit alone (slot 0, sweep 0)", 29), ("PGS slot = visit + broadcast", 30), ("noslip slot = visit + broadcast", 31)):
                v_ = t[:, col][t[:, col] > 0]
                if len(v_): add("sub0   " + nm_ + " [min over warps, median]", (v_.min(), np.median(v_) * 1965))
        add("sub0 after P4b -> stamp 6", (np.median(t[:, 6] - t[:, 26]), (t[:, 6] - t[:, 26]).max()))
print(f"N={N} warps={nw}  (cycles @1.965 GHz; median over 10 steps of [median over warps, max over warps])")
for k, v in acc.items():
    a = np.array(v)
    if "per launch" in k:                                      # event counters, not cycles
        print(f"{k:50s} {np.mean(a[:, 0]):10.3f}  {np.mean(a[:, 1]):10.3f}")
    elif a.ndim == 1:
        print(f"{k:50s} {np.median(a):10.0f} cyc  {np.median(a) / 1965:7.2f} us")
    else:
        print(f"{k:50s} {np.median(a[:, 0]):10.0f} cyc  {np.median(a[:, 0]) / 1965:7.2f} us   max-warp {np.median(a[:, 1]) / 1965:7.2f} us")

# ---- regression of the sweep phase on the contact structure of each warp (debug buffer of a debug batch)
if os.environ.get("NM_REGRESS", "1") == "1":
    from nightmare_rl_b200.batch import Batch
    cfg2 = NightmareV3Config(); cfg2.env.num_envs = N; cfg2.viewer.render = cfg2.viewer.record_states = False
    env2 = NightmareV3Env(cfg2, seed=1, device=dev, debug=True); env2.reset()
    env2.episode_length_buf = torch.randint(0, 1250, (N,), device=dev)
    for i in range(40):
        env2._batch.step(acts[i % 8], 10 + i)
    buf.zero_()
    env2._batch.step(acts[0], 100)
    torch.cuda.synchronize()
    t = buf[:nw].cpu().numpy().astype(np.float64)
    dbg = env2._batch.debug.cpu().numpy()
    ncl = dbg[:, 8:8 + 7 * 12:12].astype(int)                 # contacts per lane (0..6) of the LAST substep
    order = [6, 0, 1, 2, 3, 4, 5]
    W, S, K = [], [], []
    for w in range(nw):
        envs = ncl[4 * w:4 * w + 4]
        lists = [[e[l] for l in order if e[l] > 0] for e in envs]
        ns = max(len(x) for x in lists)
        work = sum(max((x[s] if s < len(x) else 0) for x in lists) for s in range(ns))
        W.append(work); S.append(ns); K.append(max(e.sum() for e in envs))
    W, S, K = np.array(W), np.array(S), np.array(K)
    p9 = (t[:, 18] - np.where(t[:, 17] > 0, t[:, 17], t[:, 16])) / 1965.0      # sub1: P9 + P10 [us]
    bld = (np.where(t[:, 17] > 0, t[:, 17], t[:, 16]) - t[:, 16]) / 1965.0
    col = (t[:, 16] - t[:, 15]) / 1965.0
    print("contacts per env: mean %.2f max %d; per-warp work units: mean %.1f max %d; slots mean %.1f max %d" % (ncl.sum(1).mean(), ncl.sum(1).max(), W.mean(), W.max(), S.mean(), S.max()))
    A_ = np.stack([np.ones(nw), S, W], 1)
    coef, *_ = np.linalg.lstsq(A_, p9, rcond=None)
    print("P9 time [us] ~ %.2f + %.2f * slots + %.2f * contact-units   (per sweep-set of 7); max p9 %.1f at W=%d" % (coef[0], coef[1], coef[2], p9.max(), W[p9.argmax()]))
    for wv_ in sorted(set(W.tolist())):
        m = W == wv_
        print(f"  W={wv_:2d}: warps {m.sum():4d}  P9 median {np.median(p9[m]):6.2f} us  build median {np.median(bld[m]):5.2f}  collision median {np.median(col[m]):5.2f}")
