"""The second model of the reference, models/anymal_c (BASELINE configs[3]), on the oracle: Newton solver, elliptic cones with
impratio 100 (anymal_c.xml:4), joint damping and friction loss (:9), condim-6 sphere feet with priority 1 (:20-21), position
actuators with a force range (:26), joint limits (:153...), box / cylinder / sphere against the plane, Euler integration with
implicit joint damping.  Scope statement: collisions with the floor only -- the model's geom-geom self collisions
(every collision geom has contype = conaffinity = 1) are not generated, as SURVEY.md hard part 7 scopes it.

MuJoCo itself cannot be run here; these are the invariants a correct restatement must satisfy."""
import numpy as np
import pytest

from conftest import ROOT
from nightmare_rl_b200 import mjcf
from oracle import oracle as O
import os

NMB = os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb")


@pytest.fixture(scope="module")
def cm():
    return mjcf.CompiledModel.load(NMB)


@pytest.fixture(scope="module")
def om():
    return O.OracleModel(NMB)


def test_stands_on_four_feet(cm, om):
    """ctrl = 0 holds the joints at the straight-leg pose: the robot stands on its four sphere feet (condim 6), the normal forces
    add up to its weight, and the soft feet (solimp 0.015 1 0.031) sink about two centimetres."""
    b = O.OracleBatch(om, 1)
    b.physics_step(np.zeros((1, 12)), 1500)                       # 3 s at the model's 2 ms step
    q, v, _ = b.get_state()
    con = b.get(0, "contact").reshape(-1, 7)
    assert [int(g) for g in con[:, 1]] == [20, 28, 36, 44]         # the four foot spheres, in geom order
    assert (cm.arrays["geom_type"][[20, 28, 36, 44]] == 2).all() and (cm.arrays["geom_condim"][[20, 28, 36, 44]] == 6).all()
    assert int(b.get(0, "nefc")[0]) == 12 + 4 * 6                  # 12 friction-loss rows, then 6 rows per foot
    f = b.get(0, "efc_force")
    weight = cm.arrays["body_mass"].sum() * 9.81
    assert abs(f[12::6].sum() - weight) < 1e-3 * weight
    assert np.abs(v).max() < 5e-3 and 0.59 < q[0, 2] < 0.62
    assert (-con[:, 3] > 0.01).all() and (-con[:, 3] < 0.03).all()


def _tumbling(cm, n, seed):
    rng = np.random.default_rng(seed)
    q = np.tile(cm.qpos0, (n, 1))
    q[:, 2] = rng.uniform(0.3, 0.8, n)
    q[:, 3:7] = rng.normal(size=(n, 4))
    q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
    q[:, 7:] += rng.uniform(-0.6, 0.6, (n, 12))
    v = rng.normal(size=(n, 18)) * 0.5
    return q, v, rng


def test_newton_solution_satisfies_the_optimality_conditions(cm, om):
    """At the solver's answer the gradient of MuJoCo's convex cost vanishes, M (qacc - qacc_smooth) = J' f, and every force
    obeys its constraint type: friction loss |f| <= 0.1, limits and contact normals f >= 0, contact friction inside the
    elliptic cone |f_j / mu_j| <= f_n (on it while sliding)."""
    n = 48
    q, v, rng = _tumbling(cm, n, 0)
    b = O.OracleBatch(om, n)
    b.set_state(q, v, np.zeros((n, 18)))
    ctrl = rng.uniform(-1, 1, (n, 12))
    seen_types, on_cone, worst = set(), 0, 0.0
    for it in range(30):
        b.physics_step(ctrl, 10, 8)
        b.forward(ctrl, 8)
        for i in range(n):
            ne = int(b.get(i, "nefc")[0])
            J = b.get(i, "efc_J").reshape(ne, 18)
            f = b.get(i, "efc_force")
            M = b.get(i, "M").reshape(18, 18)
            r = M @ (b.get(i, "qacc") - b.get(i, "qacc_smooth")) - J.T @ f
            worst = max(worst, np.abs(r).max() / max(1.0, np.abs(J.T @ f).max()))
            assert (np.abs(f[:12]) <= 0.1 + 1e-12).all()                                # friction loss rows come first
            con = b.get(i, "contact").reshape(-1, 7)
            row = ne - sum(int(cm.arrays["geom_condim"][int(g)]) if int(cm.arrays["geom_priority"][int(g)]) > 0 else 3 for g in con[:, 1])
            assert (f[12:row] >= 0).all()                                                # joint-limit rows
            for g in con[:, 1]:
                g = int(g)
                seen_types.add(int(cm.arrays["geom_type"][g]))
                foot = int(cm.arrays["geom_priority"][g]) > 0
                dim = 6 if foot else 3
                fr = np.array([0.8, 0.8, 0.02, 0.01, 0.01]) if foot else np.array([1.0, 1.0])
                fc = f[row:row + dim]
                assert fc[0] >= -1e-12
                t = np.linalg.norm(fc[1:] / fr[:dim - 1])
                assert t <= fc[0] * (1 + 1e-7) + 1e-9
                on_cone += t > fc[0] * (1 - 1e-6) and fc[0] > 1e-6
                row += dim
            assert row == ne
    assert worst < 1e-8, worst
    assert seen_types == {2, 5, 6} and on_cone > 50                                      # sphere, cylinder and box contacts; sliding ones


def test_joint_limits_and_force_range(cm, om):
    """In the air, a hip-abduction target far outside the joint range (-0.72, 0.49): the actuator saturates at its force range
    (kp * error = 151 > 80), the limit row switches on at the range and the joint comes to rest a little past it."""
    b = O.OracleBatch(om, 1)
    q = np.tile(cm.qpos0, (1, 1))
    q[0, 2] = 50.0
    b.set_state(q, np.zeros((1, 18)), np.zeros((1, 18)))
    ctrl = np.zeros((1, 12))
    ctrl[0, 0] = 2.0
    b.physics_step(ctrl, 1, 1)
    assert abs(b.get(0, "qfrc_actuator")[6] - 80.0) < 1e-12                              # clamped to forcerange
    b.physics_step(ctrl, 1500)
    q, v, _ = b.get_state()
    assert 0.49 < q[0, 7] < 0.56 and abs(v[0, 6]) < 1e-3
    b.forward(ctrl)
    ne = int(b.get(0, "nefc")[0])
    assert ne == 13                                                                      # 12 friction-loss rows + 1 limit row
    J = b.get(0, "efc_J").reshape(ne, 18)
    assert J[12, 6] == -1 and np.count_nonzero(J[12]) == 1                               # upper limit: d(distance)/dq = -1
    f = b.get(0, "efc_force")
    assert f[12] > 50                                                                    # holds against the saturated actuator (damping and gravity share the rest)


def test_fp32_build_agrees_on_one_step(cm):
    """The float build of the same source (what an fp32 CUDA implementation can reach) against the fp64 build, one substep from
    identical tumbling states: the rounding floor the anymal_c kernel is measured against."""
    n = 96
    q, v, rng = _tumbling(cm, n, 3)
    ctrl = rng.uniform(-1, 1, (n, 12))
    a = O.OracleBatch(O.OracleModel(NMB), n)
    f32 = O.OracleBatch(O.OracleModel(NMB, variant="f32"), n)
    a.set_state(q, v, np.zeros((n, 18)))
    errs = []
    for t in range(40):
        qa, va, wa = a.get_state()
        q32, v32, w32 = qa.astype(np.float32), va.astype(np.float32), wa.astype(np.float32)
        a.set_state(q32, v32, w32)
        f32.set_state(q32, v32, w32)
        a.physics_step(ctrl, 1, 8)
        f32.physics_step(ctrl, 1, 8)
        same = np.array([a.get(i, "ncon")[0] == f32.get(i, "ncon")[0] for i in range(n)])
        va2, vf2 = a.get_state()[1], f32.get_state()[1]
        errs.append((np.abs(va2 - vf2).max(axis=1) / np.maximum(np.abs(va2).max(axis=1), 1e-3))[same])
    e = np.concatenate(errs)
    print(f"\n[anymal f32 vs f64] one-step qvel rel: median {np.median(e):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e} over {e.size} env-substeps")
    assert np.median(e) < 1e-5 and np.percentile(e, 90) < 5e-4 and np.percentile(e, 99) < 1e-2      # elliptic cones at impratio 100 are stiff: a heavy tail
