"""Physical invariants of the CPU oracle (oracle/nm_oracle.c).  The reference pins nothing for this
path (no tests, no golden vectors; MuJoCo itself is not installable here), so the oracle is anchored on
first principles instead — SURVEY.md §4b."""
import numpy as np
import pytest

from nightmare_rl_b200 import mjcf
from oracle import oracle as O


def _variant(compiled_model, tmp_path, timestep=None, integrator=None, no_actuators=False):
    cm = mjcf.CompiledModel(dict((k, v.copy()) for k, v in compiled_model.arrays.items()), compiled_model.names)
    if no_actuators:
        cm.arrays["act_gain"][:] = 0
        cm.arrays["act_bias"][:] = 0
    if timestep is not None:
        cm.arrays["opt_real"][0] = timestep
    if integrator is not None:
        cm.arrays["opt_int"][0] = integrator
    p = str(tmp_path / f"v_{timestep}_{integrator}_{no_actuators}.nmb")
    cm.save(p)
    return cm, O.OracleModel(p)


def _mul_inert(i, v):
    r = np.zeros(6)
    r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5]
    r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5]
    r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4]
    r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3]
    r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4]
    r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5]
    return r


def _random_flying_state(cm, rng):
    qpos = cm.qpos0.copy()
    qpos[2] = 5.0
    qpos[3:7] = rng.normal(size=4)
    qpos[3:7] /= np.linalg.norm(qpos[3:7])
    qpos[7:] = rng.uniform(-0.5, 0.5, 18)
    qvel = rng.normal(size=24) * np.r_[np.ones(3) * 0.5, np.ones(3) * 2, np.ones(18) * 3]
    return qpos, qvel


def test_free_fall_law(oracle_model):
    """Contact-free, ctrl=0: v_n = -g h n and z_n = z0 - g h^2 n(n+1)/2 (semi-implicit), joints stay at rest."""
    b = O.OracleBatch(oracle_model, 1)
    h, g = 0.008, 9.81
    for n in range(1, 19):
        b.physics_step(np.zeros((1, 18)), 1)
        q, v, _ = b.get_state()
        assert b.get(0, "ncon")[0] == 0
        assert abs(v[0, 2] + g * h * n) < 1e-12
        assert abs(q[0, 2] - (0.15 - 0.5 * g * h * h * n * (n + 1))) < 1e-12
        assert np.abs(q[0, 7:]).max() < 1e-12 and np.abs(q[0, 3:7] - [1, 0, 0, 0]).max() < 1e-12
    b.physics_step(np.zeros((1, 18)), 3)     # ncon belongs to the forward pass at the START of a substep
    assert b.get(0, "ncon")[0] >= 1          # first touch-down ~0.16 s after release


def test_mass_matrix_matches_kinetic_energy(compiled_model, oracle_model):
    """q'Mq/2 equals the sum of body kinetic energies computed from the c-frame velocities."""
    rng = np.random.default_rng(1)
    b = O.OracleBatch(oracle_model, 1)
    qpos, qvel = _random_flying_state(compiled_model, rng)
    b.set_state(qpos[None], qvel[None], np.zeros((1, 24)))
    b.forward(np.zeros((1, 18)))
    M = b.get(0, "M").reshape(24, 24)
    ci, cv = b.get(0, "cinert").reshape(-1, 10), b.get(0, "cvel").reshape(-1, 6)
    T = sum(0.5 * cv[i] @ _mul_inert(ci[i], cv[i]) for i in range(1, 20))
    assert abs(0.5 * qvel @ M @ qvel - T) < 1e-10 * max(1.0, T)
    assert np.allclose(M, M.T, atol=1e-14) and np.linalg.eigvalsh(M).min() > 0


def test_forward_inverse_roundtrip(compiled_model, oracle_model):
    """M qacc_smooth + bias = actuator force."""
    rng = np.random.default_rng(2)
    b = O.OracleBatch(oracle_model, 1)
    qpos, qvel = _random_flying_state(compiled_model, rng)
    ctrl = rng.uniform(-10, 10, 18)
    b.set_state(qpos[None], qvel[None], np.zeros((1, 24)))
    b.forward(ctrl[None])
    M = b.get(0, "M").reshape(24, 24)
    res = M @ b.get(0, "qacc_smooth") + b.get(0, "qfrc_bias") - b.get(0, "qfrc_actuator")
    assert np.abs(res).max() < 1e-9
    tau = 0.8 * (np.clip(ctrl, -8, 8) - qvel[6:])                  # kv (clamp(ctrl) - qvel), mjmodel.xml:136
    assert np.allclose(b.get(0, "qfrc_actuator")[6:], tau, atol=1e-12)
    assert np.abs(b.get(0, "qfrc_actuator")[:6]).max() == 0


@pytest.mark.parametrize("h", [1e-4, 5e-5])
def test_energy_and_momentum_conservation(compiled_model, tmp_path, h):
    """Torque-free flight: energy drift is O(h) (halves with h), linear momentum follows m g t, angular
    momentum about the COM is conserved.  Exercises CRBA, RNE (incl. free-joint quasi-velocities) and
    the quaternion integrator together."""
    cm, om = _variant(compiled_model, tmp_path, timestep=h, integrator=mjcf.INT_EULER, no_actuators=True)
    rng = np.random.default_rng(0)
    b = O.OracleBatch(om, 1)
    qpos, qvel = _random_flying_state(cm, rng)
    b.set_state(qpos[None], qvel[None], np.zeros((1, 24)))

    def diag():
        b.forward(np.zeros((1, 18)))
        M = b.get(0, "M").reshape(24, 24)
        qv = b.get(0, "qvel")
        xipos = b.get(0, "xipos").reshape(-1, 3)
        ci, cv = b.get(0, "cinert").reshape(-1, 10), b.get(0, "cvel").reshape(-1, 6)
        mom = sum(_mul_inert(ci[i], cv[i]) for i in range(1, 20))
        return 0.5 * qv @ M @ qv + 9.81 * (cm.body_mass * xipos[:, 2]).sum(), mom

    e0, m0 = diag()
    T = 0.05
    b.physics_step(np.zeros((1, 18)), int(round(T / h)))
    e1, m1 = diag()
    assert abs(e1 - e0) < 8.0 * h * abs(e0)          # first-order drift, ~-1.4e-3 J at h=1e-4 on 147 J
    assert np.abs(m1[:3] - m0[:3]).max() < 0.2 * h * 1e3 * 1e-2 + 1e-5
    assert np.abs(m1[3:] - m0[3:] - np.array([0, 0, -3.0 * 9.81 * T])).max() < 1e-4


def test_energy_drift_is_first_order(compiled_model, tmp_path):
    drift = []
    for h in (1e-4, 5e-5):
        cm, om = _variant(compiled_model, tmp_path, timestep=h, integrator=mjcf.INT_EULER, no_actuators=True)
        rng = np.random.default_rng(0)
        b = O.OracleBatch(om, 1)
        qpos, qvel = _random_flying_state(cm, rng)
        b.set_state(qpos[None], qvel[None], np.zeros((1, 24)))

        def energy():
            b.forward(np.zeros((1, 18)))
            M = b.get(0, "M").reshape(24, 24)
            qv = b.get(0, "qvel")
            return 0.5 * qv @ M @ qv + 9.81 * (cm.body_mass * b.get(0, "xipos").reshape(-1, 3)[:, 2]).sum()

        e0 = energy()
        b.physics_step(np.zeros((1, 18)), int(round(0.05 / h)))
        drift.append(energy() - e0)
    assert abs(drift[0] / drift[1] - 2.0) < 0.05


def _settle(oracle_model, n_sub=120, seed=0, ctrl_scale=0.0):
    rng = np.random.default_rng(seed)
    b = O.OracleBatch(oracle_model, 1)
    for t in range(n_sub):
        b.physics_step(rng.uniform(-1, 1, (1, 18)) * ctrl_scale, 1)
    return b


def test_contact_solver_invariants(oracle_model):
    """A = J M^-1 J' + R is symmetric PSD, pyramid forces are >= 0, touch sensors add up to the normal forces."""
    seen = 0
    for seed, scale in ((0, 0.0), (1, 4.0), (2, 8.0)):
        b = _settle(oracle_model, 60 + 20 * seed, seed, scale)
        for _ in range(30):
            b.physics_step(np.zeros((1, 18)), 1)
            ne = int(b.get(0, "nefc")[0])
            if ne == 0:
                continue
            seen += 1
            A = b.get(0, "efc_AR").reshape(ne, ne)
            assert np.allclose(A, A.T, atol=1e-9 * np.abs(A).max())
            assert np.linalg.eigvalsh(0.5 * (A + A.T)).min() > 0
            f = b.get(0, "efc_force")
            assert (f >= 0).all()
            con = b.get(0, "contact").reshape(-1, 7)
            fn = f.reshape(-1, 4).sum(axis=1)
            sd = b.get(0, "sensordata")
            geoms = con[:, 1].astype(int)
            for k in range(6):                       # tibia sites have a 10 m radius: every contact of that body counts
                assert abs(sd[k] - fn[geoms == 2 + k].sum()) < 1e-9
            assert abs(sd[12] - fn[geoms == 1].sum()) < 1e-9
            assert (sd[6:12] <= sd[:6] + 1e-12).all()        # foot sphere lies inside the tibia sphere
            # qacc is consistent with the forces: M (qacc - qacc_smooth) = J' f
            M = b.get(0, "M").reshape(24, 24)
            J = b.get(0, "efc_J").reshape(ne, 24)
            assert np.abs(M @ (b.get(0, "qacc") - b.get(0, "qacc_smooth")) - J.T @ f).max() < 1e-8
            assert (con[:, 3] <= 0).all()            # active contacts penetrate (margin 0)
    assert seen > 20


def test_contact_rows_are_pyramidal(oracle_model):
    b = _settle(oracle_model, 40)
    ne = int(b.get(0, "nefc")[0])
    assert ne > 0 and ne % 4 == 0
    J = b.get(0, "efc_J").reshape(ne, 24)
    for c in range(ne // 4):
        r = J[4 * c:4 * c + 4]
        # opposite edges share the normal row: (r0 + r1)/2 == (r2 + r3)/2 == J_normal; mu = 1
        assert np.allclose(r[0] + r[1], r[2] + r[3], atol=1e-12)
        jn = 0.5 * (r[0] + r[1])
        assert np.allclose(jn[:3], [0, 0, 1], atol=1e-12)           # flat floor: normal = +z on the base translation dofs
        R = b.get(0, "efc_R")[4 * c:4 * c + 4]
        assert np.allclose(R, R[0]) and R[0] > 0


def test_rest_pose_is_stable(oracle_model):
    """Zero actions for 3 s: the robot settles on the floor (belly or feet) and stops moving."""
    b = O.OracleBatch(oracle_model, 1)
    b.physics_step(np.zeros((1, 18)), 375)
    q, v, _ = b.get_state()
    assert np.isfinite(q).all() and np.abs(v).max() < 0.2      # 3-sweep PGS leaves a small residual jitter
    assert 0.0 < q[0, 2] < 0.1
    sd = b.get(0, "sensordata")
    assert abs(sd[:6].sum() + sd[12] - 3.0 * 9.81) < 0.5            # contact forces carry the weight


def test_divergence_guard(compiled_model, oracle_model):
    b = O.OracleBatch(oracle_model, 1)
    qpos = compiled_model.qpos0.copy()
    qpos[0] = np.nan
    b.set_state(qpos[None], None, None)
    b.physics_step(np.zeros((1, 18)), 1)
    q, v, _ = b.get_state()
    assert np.isfinite(q).all() and b.get(0, "nwarn")[0] >= 1


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): the RNG that replaces numpy's MT19937."""
    assert [hex(x) for x in O.philox4x32(0, 0, 0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = 0xFFFFFFFF
    assert [hex(x) for x in O.philox4x32(f, f, f, f, f, f)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox4x32(0xA4093822, 0x299F31D0, 0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_threads_do_not_change_results(oracle_model):
    rng = np.random.default_rng(4)
    ctrl = rng.uniform(-8, 8, (16, 18))
    out = []
    for th in (1, 4):
        b = O.OracleBatch(oracle_model, 16)
        b.physics_step(ctrl, 40, th)
        out.append(b.get_state())
    for a, c in zip(*out):
        assert np.array_equal(a, c)


def test_pgs_converges_to_the_complementarity_solution(compiled_model, tmp_path):
    """With many sweeps (and no noslip pass) the projected Gauss-Seidel iteration must land on the solution of the contact LCP
    it is a solver for:  f >= 0,  (A + R) f + b >= 0,  f . ((A + R) f + b) = 0   (pyramidal rows, Appendix A.6).
    A wrong projection, update order or residual in the sweep would stall somewhere else."""
    cm = mjcf.CompiledModel(dict((k, v.copy()) for k, v in compiled_model.arrays.items()), compiled_model.names)
    cm.arrays["opt_int"][3] = 400            # iterations
    cm.arrays["opt_int"][4] = 0              # noslip_iterations
    p = str(tmp_path / "pgs400.nmb")
    cm.save(p)
    om = O.OracleModel(p)
    rng = np.random.default_rng(4)
    checked = 0
    for trial in range(6):
        b = O.OracleBatch(om, 1)
        q, v, w = b.get_state()
        q[0, 7:] += rng.uniform(-0.3, 0.3, 18)
        q[0, 2] = rng.uniform(0.03, 0.12)
        b.set_state(q, v, w)
        for t in range(40):
            b.physics_step(rng.uniform(-8, 8, (1, 18)) if t % 5 == 0 else None, 1)
            ne = int(b.get(0, "nefc")[0])
            if ne == 0:
                continue
            AR = b.get(0, "efc_AR").reshape(ne, ne)
            bb, f = b.get(0, "efc_b"), b.get(0, "efc_force")
            r = AR @ f + bb
            scale = max(1.0, np.abs(bb).max())
            assert (f >= 0).all()
            assert r.min() > -1e-4 * scale, (trial, t, r.min())          # the sweeps stop on the cost-improvement tolerance (1e-8, scaled): residual ~ its square root
            assert np.abs(f * r).max() < 1e-4 * scale * max(1.0, f.max()), (trial, t)
            checked += 1
    assert checked > 100


def _tilted(compiled_model, tmp_path, theta_deg, along):
    """Model variant whose gravity is tilted by theta towards the horizontal direction `along` (a tilted floor, seen from the floor)."""
    cm = mjcf.CompiledModel(dict((k, v.copy()) for k, v in compiled_model.arrays.items()), compiled_model.names)
    th = np.radians(theta_deg)
    cm.arrays["opt_real"][1:4] = [9.81 * np.sin(th) * np.cos(along), 9.81 * np.sin(th) * np.sin(along), -9.81 * np.cos(th)]
    p = str(tmp_path / f"tilt_{theta_deg}_{along:.2f}.nmb")
    cm.save(p)
    return cm, p


def test_pyramidal_friction_cone_holds_and_slides(compiled_model, tmp_path):
    """Friction 1 with MuJoCo's pyramidal cone: |T1| + |T2| <= mu N in the contact frame.  A robot lying on a slope therefore
    stays put below 45 degrees when the slope runs along a tangent axis (world x on the flat floor) but already slides at
    atan(1/sqrt 2) = 35.3 degrees along the diagonal -- the anisotropy is a documented property of the pyramidal approximation
    and a sharp test of how the four edge rows are built and projected."""
    def speed_after(theta, along):
        _, p = _tilted(compiled_model, tmp_path, theta, along)
        b = O.OracleBatch(O.OracleModel(p), 1)
        b.physics_step(np.zeros((1, 18)), 250)                       # 2 s: drop, settle, then hold or slide
        _, v, _ = b.get_state()
        return float(np.linalg.norm(v[0, :2]))
    assert speed_after(40, 0.0) < 0.1 and speed_after(50, 0.0) > 2.0             # along a tangent axis: threshold 45 degrees
    assert speed_after(30, np.pi / 4) < 0.1 and speed_after(40, np.pi / 4) > 1.0   # along the diagonal: threshold 35.3 degrees


def test_plane_mesh_contact_cap_is_a_model_option(compiled_model, tmp_path):
    """How many contacts one plane-mesh pair may produce is stored in the model file (opt_int[7], default 4 = support vertex + up
    to 3 neighbours as recalled in SURVEY.md Appendix A.2).  It is an option precisely because that number could not be checked
    against MuJoCo here: a cross-check that finds a different cap fixes it by re-saving the model."""
    assert int(compiled_model.arrays["opt_int"][7]) == 4
    rng = np.random.default_rng(1)
    n = 64
    qpos = np.tile(compiled_model.qpos0, (n, 1))
    qpos[:, 7:] += rng.uniform(-0.6, 0.6, (n, 18))
    qpos[:, 2] = rng.uniform(0.03, 0.22, n)
    qpos[:, 3:7] = rng.normal(size=(n, 4))
    qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
    most = {}
    for cap in (4, 3, 1):
        cm = mjcf.CompiledModel(dict((k, v.copy()) for k, v in compiled_model.arrays.items()), compiled_model.names)
        cm.arrays["opt_int"][7] = cap
        p = str(tmp_path / f"cap{cap}.nmb")
        cm.save(p)
        b = O.OracleBatch(O.OracleModel(p), n)
        b.set_state(qpos, np.zeros((n, 24)), np.zeros((n, 24)))
        most[cap] = 0
        for _ in range(60):                                      # tumbling robots land on hull faces and tibia flanks
            b.physics_step(np.zeros((n, 18)), 1, 8)
            for i in range(n):
                con = b.get(i, "contact").reshape(-1, 7)
                con = con[con[:, 0] == 0]                        # plane-mesh pairs only (tibia-tibia pairs give one contact each)
                if len(con):
                    most[cap] = max(most[cap], int(np.bincount(con[:, 1].astype(int)).max()))
        assert np.isfinite(b.get_state()[0]).all()
    assert most == {4: 4, 3: 3, 1: 1}, most


def test_recalled_details_are_model_options(compiled_model, tmp_path):
    """The other engine details SURVEY.md Appendix A could only recall -- candidate set of the extra plane-mesh contacts, how
    their separation is measured and how large it must be, the R factor of pyramidal rows, where qacc_warmstart is saved --
    are model data too (opt_int[9..11], opt_real[9..10]; nightmare_rl_b200/mjcf.py).  Each default is the recalled value, each
    alternative changes the simulation in the way its meaning says, and none needs a code change."""
    oi, orl = compiled_model.arrays["opt_int"], compiled_model.arrays["opt_real"]
    assert list(oi[8:12]) == [50, 0, 0, 0] and list(orl[8:11]) == [1e-6, 0.3, 2.0]
    rng = np.random.default_rng(1)
    n = 96
    qpos = np.tile(compiled_model.qpos0, (n, 1))
    qpos[:, 7:] += rng.uniform(-0.6, 0.6, (n, 18))
    qpos[:, 2] = rng.uniform(0.03, 0.22, n)
    qpos[:, 3:7] = rng.normal(size=(n, 4))
    qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)

    def run(name, key=None, idx=None, val=None, steps=30):
        cm = mjcf.CompiledModel(dict((k, v.copy()) for k, v in compiled_model.arrays.items()), compiled_model.names)
        if key is not None:
            cm.arrays[key][idx] = val
        p = str(tmp_path / f"{name}.nmb")
        cm.save(p)
        b = O.OracleBatch(O.OracleModel(p), n)
        b.set_state(qpos, np.zeros((n, 24)), np.zeros((n, 24)))
        ncon, first = 0, None
        for s in range(steps):
            b.physics_step(np.zeros((n, 18)), 1, 8)
            cnt = np.array([b.get(i, "ncon")[0] for i in range(n)])
            ncon += int(cnt.sum())
            if s == 0:
                first = dict(ncon=cnt, R=[b.get(i, "efc_R") for i in range(n)], qacc=[b.get(i, "qacc") for i in range(n)], warm=b.get_state()[2].copy())
        q, v, w = b.get_state()
        assert np.isfinite(q).all() and np.isfinite(v).all()
        return dict(q=q, ncon=ncon, first=first)

    base = run("default")
    allv = run("allverts", "opt_int", 9, 1)
    # more candidates can only add contacts on the first substep (same state, same support vertices)
    assert (allv["first"]["ncon"] >= base["first"]["ncon"]).all() and (allv["first"]["ncon"] > base["first"]["ncon"]).any()
    sepv = run("sepvert", "opt_int", 10, 1)
    sep = run("sep015", "opt_real", 9, 0.15)
    assert (sep["first"]["ncon"] >= base["first"]["ncon"]).all() and (sep["first"]["ncon"] > base["first"]["ncon"]).any()   # closer contacts allowed
    assert np.isfinite(sepv["q"]).all()                                                            # (may coincide with the default on a sample: vertex and point are dist/2 apart)
    rf = run("rfac1", "opt_real", 10, 1.0)
    ratios = np.concatenate([a / b for a, b in zip(rf["first"]["R"], base["first"]["R"]) if len(a) and len(a) == len(b)])
    assert len(ratios) > 100 and np.allclose(ratios, 0.5, rtol=1e-12)                              # R of every pyramid edge halves
    wa = run("warmafter", "opt_int", 11, 1)
    inc = [i for i in range(n) if base["first"]["ncon"][i] > 0]
    # first substep: identical solve, different save point -- after noslip the warm start IS that substep's qacc, before it is not
    assert all(np.array_equal(wa["first"]["qacc"][i], base["first"]["qacc"][i]) for i in inc)
    assert all(np.array_equal(wa["first"]["warm"][i], wa["first"]["qacc"][i]) for i in inc)
    assert any(not np.array_equal(base["first"]["warm"][i], base["first"]["qacc"][i]) for i in inc)


def test_passive_contact_never_gains_energy(compiled_model, oracle_model):
    """Zero controls: the velocity servos act as dampers and the soft contacts follow a critically damped reference
    (solref 0.02 / 1), so a robot dropped in any pose must never have more mechanical energy than it started with, and must end
    well below it.  (Within one impact the energy is not monotone: at dt = 8 ms the contact 'spring' takes up joules at 4 cm of
    penetration and the pyramidal edges that oppose sliding keep pushing while the hull already separates, so a tumbling robot
    bounces -- losing energy on every bounce.)  A sign error in aref, in the pyramid rows or in J'f would pump energy in."""
    rng = np.random.default_rng(9)
    n = 16
    qpos = np.tile(compiled_model.qpos0, (n, 1))
    qpos[:, 7:] += rng.uniform(-0.5, 0.5, (n, 18))
    qpos[:, 2] = rng.uniform(0.35, 0.45, n)                   # clear of the floor in every orientation: no energy pre-stored in penetration
    qpos[:, 3:7] = rng.normal(size=(n, 4))
    qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
    b = O.OracleBatch(oracle_model, n)
    b.set_state(qpos, np.zeros((n, 24)), np.zeros((n, 24)))

    def energy():
        b.forward(np.zeros((n, 18)), 8)
        out = np.zeros(n)
        for i in range(n):
            M = b.get(i, "M").reshape(24, 24)
            qv = b.get(i, "qvel")
            z = b.get(i, "xipos").reshape(-1, 3)[:, 2]
            out[i] = 0.5 * qv @ M @ qv + 9.81 * (compiled_model.body_mass * z).sum()
        return out

    e0 = energy()
    for _ in range(375):                                       # 3 s
        b.physics_step(np.zeros((n, 18)), 1, 8)
        e = energy()
        assert (e <= e0 + 0.02).all(), float((e - e0).max())
    assert (e0 - e > 6.0).all()                                # dropped from ~0.4 m: m g h = 11.8 J, ~2.5 J left lying on the floor
    assert np.abs(b.get_state()[1]).max() < 1.0                # and it has come to rest
