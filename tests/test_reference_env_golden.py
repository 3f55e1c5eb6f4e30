"""The env layer against the reference's OWN code.

`tests/golden/reference_env_on_oracle_physics.npz` holds what the unmodified `NightmareV3Env.step` of the reference
(envs/nightmare_v3_env.py:145-371,399-497) returns when its seven MuJoCo calls are served by this repository's oracle
physics (tools/make_refenv_golden.py explains the stand-in and the shared Philox command stream).  Here the oracle's C
restatement of that env layer is driven with the same actions: with identical physics underneath, every output of the
reference code must be reproduced -- flags and counters exactly, floats to fp64->fp32 rounding.  The CUDA kernel is compared
with the same fixture in tests/test_gpu_env.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import NMB, ROOT
from nightmare_rl_b200.envcfg import REWARD_TERMS, build_envcfg
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from oracle import oracle as O

FIX = os.path.join(ROOT, "tests", "golden", "reference_env_on_oracle_physics.npz")


def scenario_cfg(name, n):
    cfg = NightmareV3Config()
    cfg.env.num_envs = n
    if name == "all_terms":                                     # same edits as tools/make_refenv_golden.py::scenarios
        cfg.env.tibia_contact_mode = 2
        cfg.env.body_contact_mode = 2
        s = cfg.rewards.scales
        s.lin_vel_z, s.ang_vel_xy, s.base_height, s.torques, s.dof_vel, s.feet_air_time, s.stand_still, s.feet_contact_forces = (
            -2.0, -0.05, -1.0, -1e-5, -1e-4, 1.0, -0.5, -0.01)
    if name == "noise":
        cfg.noise.add_noise = True
    return cfg


@pytest.mark.parametrize("name", ["default", "all_terms", "noise"])
def test_oracle_env_layer_reproduces_reference_code(name):
    g = np.load(FIX)
    seed = int(g["seed"])
    acts, ep0 = g[f"{name}.actions"], g[f"{name}.ep0"]
    T, n = acts.shape[:2]
    om = O.OracleModel(NMB)
    b = O.OracleBatch(om, n, seed=seed, envcfg=build_envcfg(scenario_cfg(name, n), 0.008))
    b.env_reset_idx(np.arange(n))
    b.env_set("ep_len", ep0.astype(np.float64))
    keys = [str(k) for k in g[f"{name}.ep_keys"]]
    worst = dict(obs=0.0, rew=0.0, qpos=0.0, cmd=0.0, ep=0.0)
    resets = 0
    for t in range(T):
        obs, rew, done, tout, means, nres = b.env_step(acts[t])
        assert np.array_equal(done, g[f"{name}.done"][t]), f"step {t}: reset flags"
        assert np.array_equal(tout.astype(bool), g[f"{name}.time_out"][t]), f"step {t}: time-out flags"
        assert np.array_equal(b.env_get("ep_len").astype(np.int64), g[f"{name}.ep_len"][t]), f"step {t}: episode lengths"
        worst["cmd"] = max(worst["cmd"], float(np.abs(b.env_get("commands") - g[f"{name}.commands"][t]).max()))
        worst["obs"] = max(worst["obs"], float(np.abs(obs - g[f"{name}.obs"][t]).max()))
        worst["rew"] = max(worst["rew"], float(np.abs(rew - g[f"{name}.rew"][t]).max()))
        worst["qpos"] = max(worst["qpos"], float(np.abs(b.get_state()[0] - g[f"{name}.qpos"][t]).max()))
        resets += int(done.sum())
        if nres:                                                # extras["episode"] is refreshed on steps with >= 1 reset (quirk Q10)
            ref = dict(zip(keys, g[f"{name}.ep_vals"][t]))
            for k, term in enumerate(REWARD_TERMS):
                if "rew_" + term in ref:
                    worst["ep"] = max(worst["ep"], abs(float(ref["rew_" + term]) - means[k]))
            # time_outs in the extras are the flags of THIS step
            assert np.array_equal(g[f"{name}.time_outs_extra"][t] > 0, tout > 0)
    print(f"\n[reference env code, {name}] {T} steps x {n} envs, {resets} resets: worst |obs| {worst['obs']:.1e} |rew| {worst['rew']:.1e} "
          f"|qpos| {worst['qpos']:.1e} |commands| {worst['cmd']:.1e} |episode means| {worst['ep']:.1e}")
    assert resets >= 4
    assert worst["cmd"] < 1e-6 and worst["qpos"] < 1e-6
    assert worst["obs"] < 1e-5 and worst["rew"] < 1e-6 and worst["ep"] < 1e-6


def test_reward_term_set_matches_reference_extras():
    """The keys the reference puts into extras['episode'] are exactly the non-zero scales (+ termination), default config."""
    g = np.load(FIX)
    assert [str(k) for k in g["default.ep_keys"]] == ["rew_" + k for k in ("action_rate", "body_contact_forces", "default_position", "dof_acc",
                                                                           "orientation", "termination", "tracking_ang_vel", "tracking_lin_vel")]


@pytest.mark.skipif(not os.path.isdir("/root/reference/envs"), reason="reference tree not present (GPU box)")
def test_fixture_regenerates_from_reference_tree(tmp_path):
    """Re-run the reference code over the oracle and compare with the committed fixture (oracle or generator drift shows here)."""
    src = open(os.path.join(ROOT, "tools", "make_refenv_golden.py")).read()
    src = src.replace('path = os.path.join(ROOT, "tests", "golden", "reference_env_on_oracle_physics.npz")', f'path = r"{tmp_path}/out.npz"')
    src = src.replace('ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', f'ROOT = r"{ROOT}"')
    (tmp_path / "gen.py").write_text(src)
    subprocess.run([sys.executable, str(tmp_path / "gen.py")], check=True, capture_output=True, timeout=600)
    a, b = np.load(FIX), np.load(tmp_path / "out.npz")
    assert set(a.files) == set(b.files)
    for k in a.files:
        if a[k].dtype.kind in "fc":
            assert np.allclose(a[k], b[k], rtol=0, atol=1e-6), k
        else:
            assert np.array_equal(a[k], b[k]), k


def test_reset_reproduces_reference_code():
    """`reset()` of the reference = reset_idx(all) followed by one zero-action step whose observations it returns (:392-396);
    the steps after it see the stale-buffer quirks of a just-reset batch (Q2: first PD law from the buffers of the zero step)."""
    g = np.load(FIX)
    acts = g["via_reset.actions"]
    T, n = acts.shape[:2]
    om = O.OracleModel(NMB)
    b = O.OracleBatch(om, n, seed=int(g["seed"]), envcfg=build_envcfg(scenario_cfg("via_reset", n), 0.008))
    b.env_reset_idx(np.arange(n))
    obs0 = b.env_step(np.zeros((n, 18), dtype=np.float32))[0]
    assert np.array_equal(obs0, g["via_reset.reset_obs"])
    for t in range(T):
        obs, rew, done, tout, _, _ = b.env_step(acts[t])
        assert np.array_equal(obs, g["via_reset.obs"][t]) and np.array_equal(rew, g["via_reset.rew"][t])
        assert np.array_equal(done, g["via_reset.done"][t])
        assert np.array_equal(b.env_get("ep_len").astype(np.int64), g["via_reset.ep_len"][t])
