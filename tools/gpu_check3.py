import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import gpu_common as G
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
n, T = 256, 50
cfg, ob, gb = G.make_env_pair(n, 1, NightmareV3Config())
ob.env_reset_idx(np.arange(n))
rng = np.random.default_rng(1)
ob.env_set("ep_len", rng.integers(0, 1251, n).astype(np.float64))
er, eo, rr = [], [], []
for t in range(T):
    G.sync_env_from_oracle(ob, gb)
    a = rng.normal(size=(n, 18)).astype(np.float32)
    obs, rew, done, tout, means, nres = ob.env_step(a)
    gb.step(torch.from_numpy(a), t + 1); torch.cuda.synchronize()
    er.append(np.abs(gb.rew.cpu().numpy() - rew)); rr.append(np.abs(rew)); eo.append(np.abs(gb.obs.cpu().numpy() - obs).max(1))
    if t % 10 == 0:
        i = int(er[-1].argmax()); print(t, "rew err max", er[-1].max(), "rew", rew[i], "obs err max", eo[-1].max(), "argmax col", np.abs(gb.obs.cpu().numpy() - obs)[int(eo[-1].argmax())].argmax())
er, eo, rr = np.concatenate(er), np.concatenate(eo), np.concatenate(rr)
print("rew abs err: median %.2e p99 %.2e max %.2e | rel to |rew|: median %.2e p99 %.2e max %.2e" % (np.median(er), np.percentile(er, 99), er.max(), np.median(er / np.maximum(rr, 1e-2)), np.percentile(er / np.maximum(rr, 1e-2), 99), (er / np.maximum(rr, 1e-2)).max()))
print("obs abs err: median %.2e p99 %.2e max %.2e" % (np.median(eo), np.percentile(eo, 99), eo.max()))
