"""Scripted gait of the reference (nikengine, custom_play.py:49-74) as a deterministic walking workload: the committed
fixture is pinned against the reference's own code when the reference tree is present, and the oracle's physics is checked
for the behaviour the gait is designed to produce (stand-up, alternating tripod, forward / backward / turning motion)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import NMB, ROOT
from oracle import oracle as O

FIX = os.path.join(ROOT, "tests", "golden", "nikengine_gait_targets.npz")


def _yaw(q):
    w, x, y, z = q
    return np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))


@pytest.mark.skipif(not os.path.isdir("/root/reference/nikengine"), reason="reference tree not present (GPU box)")
def test_fixture_reproduces_from_reference_engine(tmp_path):
    """The fixture is the reference gait engine's output, bit for bit (generator: tools/make_gait_golden.py)."""
    # run the generator into a scratch root so the committed file is not touched
    scratch = tmp_path / "tests" / "golden"
    scratch.mkdir(parents=True)
    src = open(os.path.join(ROOT, "tools", "make_gait_golden.py")).read().replace(
        'ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', f'ROOT = r"{tmp_path}"')
    (tmp_path / "gen.py").write_text(src)
    subprocess.run([sys.executable, str(tmp_path / "gen.py")], check=True, capture_output=True, timeout=300)
    a, b = np.load(FIX), np.load(scratch / "nikengine_gait_targets.npz")
    for k in ("raw", "targets", "commands"):
        assert np.array_equal(a[k], b[k]), k


def test_fixture_shape_and_limits():
    z = np.load(FIX)
    T = z["targets"].shape[0]
    assert z["targets"].shape == (T, 18) and z["raw"].shape == (T, 18) and z["commands"].shape == (T, 3)
    assert np.abs(np.diff(z["targets"], axis=0)).max() <= float(z["rate"]) + 1e-12          # custom_play.py:18,73 rate limit
    assert np.abs(z["targets"] + np.array([0, np.pi / 5, 0] * 6)).max() < 1.0               # representable as env actions


def _play_oracle(kp=12.0):
    """custom_play.py:73-76: ctrl = (target - qpos[-18:]) * kp; mj_step(model, data, 2)."""
    z = np.load(FIX)
    om = O.OracleModel(NMB)
    b = O.OracleBatch(om, 1)
    log = []
    for th in z["targets"]:
        q, _, _ = b.get_state()
        b.physics_step(((th - q[0, -18:]) * kp)[None], 2, 1)
        q, v, _ = b.get_state()
        s = b.get(0, "sensordata")
        log.append(np.r_[q[0, :7], s[6:12] > 0, s[12]])
    return z, np.array(log)


def test_oracle_walks_with_the_reference_gait():
    z, log = _play_oracle()
    cmd = z["commands"]
    dt = float(z["dt"])
    assert np.isfinite(log).all()
    # 1. the engine's wake-up sequence lifts the hull off the floor and keeps it level
    stand = log[260:300]
    assert 0.07 < stand[:, 2].mean() < 0.11 and (stand[:, 13] == 0).all() and (stand[:, 7:13].sum(1) >= 3).all()
    # 2. walking segments: motion along the body's forward axis (-x of base_link, nikengine's convention) with the sign of the
    #    command, equal speed forwards and backwards, no hull contact, hull height held
    def seg(a, b):
        y0 = _yaw(log[a, 3:7])
        d = log[b - 1, :2] - log[a, :2]
        return -(np.cos(y0) * d[0] + np.sin(y0) * d[1]) / ((b - a - 1) * dt), (_yaw(log[b - 1, 3:7]) - y0) / ((b - a - 1) * dt)
    fwd = np.flatnonzero((cmd[:, 0] > 0) & (cmd[:, 2] == 0))
    back = np.flatnonzero(cmd[:, 0] < 0)
    turn = np.flatnonzero(cmd[:, 2] > 0)
    v_f, w_f = seg(fwd[0] + 60, fwd[-1] + 1)            # skip the gait's start-up transient
    v_b, w_b = seg(back[0] + 40, back[-1] + 1)
    v_t, w_t = seg(turn[0] + 40, turn[0] + 100)
    assert 0.08 < v_f < 0.25 and -0.25 < v_b < -0.08 and abs(v_f + v_b) < 0.06, (v_f, v_b)
    assert abs(w_f) < 0.08 and abs(w_b) < 0.08 and 0.25 < w_t < 0.8, (w_f, w_b, w_t)
    walk = log[fwd[0] + 60: back[-1] + 1]
    assert (walk[:, 13] == 0).all() and 0.07 < walk[:, 2].min() and walk[:, 2].max() < 0.11
    # 3. alternating tripod: legs {0,2,4} and {1,3,5} are in stance in anti-phase
    c = walk[:, 7:13].astype(float)
    a, bb = c[:, [0, 2, 4]].mean(1), c[:, [1, 3, 5]].mean(1)
    assert np.corrcoef(a, bb)[0, 1] < -0.5
    for tri in ([0, 2, 4], [1, 3, 5]):
        assert np.corrcoef(c[:, tri].T).min() > 0.5
    # at every instant at least one tripod carries the robot
    assert (np.maximum(c[:, [0, 2, 4]].sum(1), c[:, [1, 3, 5]].sum(1)) >= 2).mean() > 0.97


def test_scripted_gait_action_source_on_cpu():
    """ScriptedGait (product-side helper) is plain torch: index arithmetic, phase shifts, looping and the action map
    a = (theta + default_dof_pos) / action_scale can be checked without a GPU."""
    import torch
    from nightmare_rl_b200.envs.scripted_gait import ScriptedGait
    z = np.load(FIX)
    g = ScriptedGait(FIX, 5, torch.device("cpu"), action_scale=0.2, phase_shift=11)
    assert g.index(0).tolist() == [0, 11, 22, 33, 44]
    a0 = g.actions()                                             # internal counter: step 0
    default = np.array([0.0, np.pi / 5, 0.0] * 6)
    assert np.allclose(a0[2].numpy() * 0.2 - default, z["targets"][22], atol=1e-6)
    assert g.t == 1 and np.allclose(g.joint_targets(3)[1].numpy(), z["targets"][14], atol=1e-7)
    lo, hi = g.loop
    assert z["commands"][lo].any() and z["commands"][hi - 1].any() and not z["commands"][lo - 1].any()
    far = g.index(10 * len(z["targets"]) + 7)
    assert ((far >= lo) & (far < hi)).all() and len(set(far.tolist())) == 5      # loops inside the walking part, envs stay out of phase
    assert (a0.abs() * 0.2 <= 1.0).all()                          # inside the env's clip range
