"""The oracle's C env layer (oracle/nm_oracle.c: env_step_one) against a statement-by-statement numpy
transcription of the reference env (tests/ref_mirror.py ≙ envs/nightmare_v3_env.py:145-371)."""
import numpy as np
import pytest

from nightmare_rl_b200.envcfg import REWARD_TERMS, build_envcfg, reward_table
from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
from oracle import oracle as O
from ref_mirror import MirrorEnv


def _run(cfg, oracle_model, steps, seed, ep0=None, actions_scale=1.0):
    n = cfg.env.num_envs
    mirror = MirrorEnv(cfg, oracle_model, seed=seed)
    ob = O.OracleBatch(oracle_model, n, seed=seed, envcfg=build_envcfg(cfg, 0.008))
    if ep0 is not None:
        mirror.episode_length_buf[:] = ep0
        ob.env_set("ep_len", ep0)
    rng = np.random.default_rng(seed)
    for t in range(steps):
        a = (rng.normal(size=(n, 18)) * actions_scale).astype(np.float32)
        o1, r1, d1 = mirror.step(a)
        o2, r2, d2, tout, means, nres = ob.env_step(a)
        assert np.array_equal(d1, d2), f"step {t}: reset flags differ"
        assert np.array_equal(mirror.time_out_buf.astype(np.float32), tout)
        assert np.allclose(o1, o2, rtol=0, atol=1e-6), f"step {t}: obs differ by {np.abs(o1 - o2).max()}"
        assert np.allclose(r1, r2, rtol=1e-6, atol=1e-7), f"step {t}: reward differs"
        assert np.array_equal(mirror.episode_length_buf, ob.env_get("ep_len").astype(np.int64))
        assert np.allclose(mirror.commands, ob.env_get("commands"), atol=1e-12)
        if nres:
            for k, name in enumerate(REWARD_TERMS):
                if name in mirror.episode_sums:
                    assert abs(mirror.extras["episode"]["rew_" + name] - means[k]) < 1e-9
        for name in mirror.episode_sums:
            assert np.allclose(mirror.episode_sums[name], ob.env_get("episode_sums")[:, REWARD_TERMS.index(name)], atol=1e-9)
    return mirror, ob


def test_default_config_with_timeouts_and_resampling(oracle_model):
    cfg = NightmareV3Config()
    cfg.env.num_envs = 6
    # episode counters placed so that %625 resampling, the strict >1250 time-out and both at once all fire
    ep0 = np.array([620, 1245, 0, 1249, 624, 100], dtype=np.float64)
    mirror, _ = _run(cfg, oracle_model, 40, seed=3, ep0=ep0)
    assert mirror.common_step_counter == 40


def test_falls_and_contact_terminations(oracle_model):
    cfg = NightmareV3Config()
    cfg.env.num_envs = 4
    cfg.env.termination_contact_force = 20.0     # make foot-force terminations frequent
    cfg.env.tibia_contact_mode = 2
    cfg.env.body_contact_mode = 2
    _run(cfg, oracle_model, 60, seed=5, actions_scale=3.0)


def test_all_reward_terms_enabled(oracle_model):
    cfg = NightmareV3Config()
    cfg.env.num_envs = 3
    s = cfg.rewards.scales
    s.lin_vel_z, s.ang_vel_xy, s.feet_air_time, s.torques, s.base_height = -2.0, -5.0, -4.0, -1e-5, -2000.0
    s.feet_contact_forces, s.dof_vel, s.stand_still = -0.05, -0.001, -1.0
    _run(cfg, oracle_model, 50, seed=7)


def test_reward_table_order_and_missing_functions():
    cfg = NightmareV3Config()
    table, order, keys = reward_table(cfg, 0.016)
    assert order == ["action_rate", "body_contact_forces", "default_position", "dof_acc", "orientation", "tracking_ang_vel", "tracking_lin_vel"]
    assert keys[5] == "termination" and len(keys) == 8
    assert abs(table[REWARD_TERMS.index("termination")] + 3.2) < 1e-12
    cfg.rewards.scales.collision = -1.0          # scale without a _reward_ function: the reference raises at construction
    with pytest.raises(AttributeError):
        reward_table(cfg, 0.016)


def test_envcfg_constants():
    ec = build_envcfg(NightmareV3Config(), 0.008)
    assert ec.resample_period == 625 and ec.max_episode_length == 1250.0 and abs(ec.dt - 0.016) < 1e-15
    nv = np.array(ec.noise_vec[:])
    assert np.allclose(nv[:12], [0.2] * 3 + [0.025] * 3 + [0.1] * 3 + [0] * 3)
    assert np.allclose(nv[12:24], 0.1) and np.allclose(nv[24:36], 0.005) and np.allclose(nv[36:], 0)   # quirk Q8
