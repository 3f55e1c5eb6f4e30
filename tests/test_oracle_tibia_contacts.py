"""Tibia-tibia convex-convex contacts of the oracle (reference models/nightmare_v3/mjmodel.xml:47: tibia geoms have
contype=2 / conaffinity=3, i.e. the 15 tibia pairs collide; MuJoCo 3.1.2 uses libccd's MPR for mesh-mesh pairs).

MuJoCo cannot be run here, so the restated narrow phase is checked against geometry that can be computed independently
(the Minkowski difference of the two hulls, qhull) and against physical invariants of self-contact (internal forces)."""
import numpy as np
import pytest

from conftest import NMB
from nightmare_rl_b200 import mjcf
from oracle import oracle as O


@pytest.fixture(scope="module")
def cm():
    return mjcf.CompiledModel.load(NMB)


def _hull_world(cm, b, env, geom):
    adr, num = int(cm.arrays["geom_hull_adr"][geom]), int(cm.arrays["geom_hull_num"][geom])
    body = int(cm.arrays["geom_body"][geom])
    R = b.get(env, "xmat").reshape(-1, 3, 3)[body]
    p = b.get(env, "xpos").reshape(-1, 3)[body]
    return cm.arrays["hull_vert"][adr:adr + num].astype(np.float64) @ R.T + p


def _airborne(cm, n, rng, spread):
    q = np.tile(cm.qpos0, (n, 1))
    q[:, 2] = 1.0                                                    # far above the floor: only self-contacts can appear
    q[:, 7:] += rng.uniform(-spread, spread, (n, 18))
    return q


def test_mpr_against_the_minkowski_difference(cm):
    """Two convex hulls intersect iff the origin lies inside their Minkowski difference.  For every tibia pair whose
    bounding spheres overlap: (i) the oracle reports a contact exactly when the origin is inside hull(A - B) (qhull);
    (ii) its depth is >= the exact minimum translation distance and <= the support of the difference along its own normal
    (MPR measures penetration through the portal the centre-to-centre ray leaves by, not the global minimum); (iii) the
    point depth * normal lies ON the boundary of the difference, i.e. translating B by it brings the hulls to touching."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(0)
    n = 96
    om = O.OracleModel(NMB)
    b = O.OracleBatch(om, n)
    b.set_state(_airborne(cm, n, rng, 1.0), np.zeros((n, 24)), np.zeros((n, 24)))
    b.forward(np.zeros((n, 18)), 8)
    tib = [g for g in range(2, 8)]
    checked = hits = 0
    for env in range(n):
        con = b.get(env, "contact").reshape(-1, 7)
        nrm = b.get(env, "contact_frame").reshape(-1, 3)
        assert (con[:, 0] >= 2).all()                                # airborne: no plane contacts
        found = {(int(c[0]), int(c[1])): (-c[3], nrm[k], c[4:7]) for k, c in enumerate(con)}
        hulls = {g: _hull_world(cm, b, env, g) for g in tib}
        for i, g1 in enumerate(tib):
            for g2 in tib[i + 1:]:
                A, B = hulls[g1], hulls[g2]
                if np.linalg.norm(A.mean(0) - B.mean(0)) > 0.3:
                    assert (g1, g2) not in found
                    continue
                md = (A[:, None, :] - B[None, :, :]).reshape(-1, 3)
                hull = ConvexHull(md)
                # facet equations: n.x + d <= 0 inside; distance of the origin to each facet plane = -d
                inside_margin = -hull.equations[:, 3].max()        # > 0: origin strictly inside; its value = exact penetration depth
                checked += 1
                if abs(inside_margin) < 1e-5:
                    continue                                         # touching within tolerance: either answer is fine
                assert ((g1, g2) in found) == (inside_margin > 0), (env, g1, g2, inside_margin)
                if (g1, g2) in found:
                    hits += 1
                    depth, normal, pos = found[(g1, g2)]
                    exact = -hull.equations[:, 3].max()              # minimum translation distance
                    assert abs(np.linalg.norm(normal) - 1) < 1e-9
                    assert depth >= exact - 1e-6
                    # support of the difference along the reported normal: the boundary point the normal ray reaches
                    h = (md @ normal).max()
                    assert h >= depth - 1e-6
                    # B + depth*normal: the origin is now ON the boundary of A - (B + depth n) = md - depth n
                    eq = hull.equations
                    shifted = (eq[:, :3] @ (depth * normal)) + eq[:, 3]    # n.(0 + depth*normal) + d  for the unshifted hull
                    assert shifted.max() > -2e-6 and shifted.max() < 2e-5
                    # the contact point lies inside both hulls (it is the midpoint of the two witness points)
                    assert ((A - pos) @ normal).max() > -1e-6 and ((B - pos) @ normal).min() < 1e-6
    assert checked > 150 and hits > 15, (checked, hits)


def _contacts(variant, q):
    n = len(q)
    b = O.OracleBatch(O.OracleModel(NMB, variant=variant), n)
    b.set_state(q, np.zeros((n, 24)), np.zeros((n, 24)))
    b.forward(np.zeros((n, 18)), 8)
    out = []
    for i in range(n):
        con, nrm = b.get(i, "contact").reshape(-1, 7), b.get(i, "contact_frame").reshape(-1, 3)
        out.append({(int(c[0]), int(c[1])): (c[3], nrm[k], c[4:7]) for k, c in enumerate(con)})
    return out


def stable_pairs(q, rng, trials=3, eps=1e-12, tol=1e-8):
    """MPR on polytopes is not continuous in its inputs: whenever the portal holds two Minkowski points that share a vertex of
    one hull, the portal normal is perpendicular to an edge of the other hull and the next support query is an exact tie
    between that edge's end points -- rounding decides, and the two answers differ by up to millimetres of depth.  ~15 % of
    the intersecting tibia pairs are like that (the foot caps are finely tessellated).  No implementation can reproduce those
    without bit-identical arithmetic (that includes MuJoCo itself), so parity is asserted on the pairs whose fp64 result does
    not move under a 1e-12 rad perturbation of the joint angles -- decided by the oracle alone."""
    base = _contacts("f64", q)
    ok = [set(d.keys()) for d in base]
    for _ in range(trials):
        qp = q.astype(np.float64).copy()
        qp[:, 7:] += rng.normal(size=(len(q), 18)) * eps
        for e, d in enumerate(_contacts("f64", qp)):
            ok[e] = {k for k in ok[e] if k in d and abs(d[k][0] - base[e][k][0]) < tol}
    return base, ok


def test_fp32_build_finds_the_same_pairs(cm):
    """The float build of the same source (the arithmetic the CUDA kernel uses) agrees on WHICH pairs touch (grazing contacts
    aside) and, on the pairs the oracle itself calls stable, on depth / normal / position for at least 90 % of them (the rest: ties resolved differently by float rounding)."""
    rng = np.random.default_rng(1)
    n = 1024
    q = _airborne(cm, n, rng, 1.0).astype(np.float32)
    q[:, 2] = 0.2
    base, ok = stable_pairs(q, rng)
    f32 = _contacts("f32", q)
    pairs = stable = agree = 0
    for e in range(n):
        for key, (dist, nrm, pos) in base[e].items():
            pairs += 1
            if key not in f32[e]:
                assert -dist < 1e-4, "only a grazing contact may be missed"
                continue
            if key in ok[e]:
                stable += 1
                d2, n2, p2 = f32[e][key]
                agree += abs(d2 - dist) < 2e-6 and np.linalg.norm(n2 - nrm) < 1e-3 and np.linalg.norm(p2 - pos) < 1e-5
        for key, (dist, _, _) in f32[e].items():
            assert key in base[e] or -dist < 1e-4
    print(f"\n[mpr f32 vs f64] {pairs} intersecting pairs, {stable} stable under a 1e-12 perturbation, {agree} of those agree to 2e-6 m")
    assert pairs > 100 and stable > 0.75 * pairs and agree >= 0.9 * stable


def test_self_contact_is_an_internal_force(cm):
    """Two tibias pressed together while the robot is in free fall: the contact forces are internal, so the whole-robot
    centre of mass keeps falling at g and the total angular momentum stays zero; the rows of the contact Jacobian have no
    base-translation component; both tibias' touch sensors read the same normal force (equal and opposite)."""
    om = O.OracleModel(NMB)
    q = np.tile(cm.qpos0, (1, 1))
    q[0, 2] = 20.0
    q[0, 7 + 0] += 0.45                                              # leg 1 and leg 2 coxae turned towards each other
    q[0, 7 + 3] -= 0.45
    b = O.OracleBatch(om, 1)
    b.set_state(q, np.zeros((1, 24)), np.zeros((1, 24)))
    ctrl = np.zeros((1, 18))
    ctrl[0, 0], ctrl[0, 3] = 1.0, -1.0                              # velocity servos keep pushing them into each other (0.8 N m)
    mass = cm.arrays["body_mass"]
    touched = 0
    depth_max = 0.0
    for t in range(120):
        b.forward(ctrl)
        con = b.get(0, "contact").reshape(-1, 7)
        if len(con):
            assert set(map(tuple, con[:, :2].astype(int))) <= {(2, 3)}
            touched += 1
            depth_max = max(depth_max, float(-con[:, 3].min()))
            depth_last = float(-con[:, 3].min())
            J = b.get(0, "efc_J").reshape(-1, 24)
            assert np.abs(J[:, :3]).max() < 1e-12                    # d(gap)/d(base translation) = 0: both bodies move with the base
            assert np.abs(J[:, 3:6]).max() < 1e-9                    # ... and rotate with it
            assert np.abs(J[:, 12:]).max() == 0                      # legs 3..6 are not involved
            s = b.get(0, "sensordata")
            f = b.get(0, "efc_force").sum()
            assert f >= 0 and abs(s[0] - f) < 1e-9 and abs(s[1] - f) < 1e-9      # tibia-1 and tibia-2 sensors (r = 10 m spheres)
        b.physics_step(ctrl, 1)
        # whole-robot momentum
        xipos = b.get(0, "xipos").reshape(-1, 3)
        cvel = b.get(0, "cvel").reshape(-1, 6)
    assert touched > 60 and depth_max < 1e-2 and depth_last < 3e-3, (touched, depth_max, depth_last)     # 7 mm on impact, 1.6 mm at rest
    qq, qv = b.get_state()[0][0], b.get_state()[1][0]
    assert 0.55 < qq[7] < 0.7 and -0.7 < qq[10] < -0.55 and np.abs(qv[[6, 9]]).max() < 0.05      # the legs stopped each other; without the contact they would be at +-1.4 rad
    # free fall of the COM: after n substeps of semi-implicit Euler v = -g*h*n exactly (internal forces cancel)
    b.forward(ctrl)
    cvel = b.get(0, "cvel").reshape(-1, 6)
    xipos = b.get(0, "xipos").reshape(-1, 3)
    com = (mass[:, None] * xipos).sum(0) / mass.sum()
    # cvel is expressed at the subtree COM: linear velocity of body i's COM = v + w x (xipos_i - com)
    vcom = sum(mass[i] * (cvel[i, 3:] + np.cross(cvel[i, :3], xipos[i] - com)) for i in range(1, len(mass))) / mass.sum()
    # (exact for the generalized momentum M(q_t) dv; the COM velocity itself drifts at O(h^2) as the legs move)
    assert np.abs(vcom[:2]).max() < 1e-5 and abs(vcom[2] + 9.81 * 0.008 * 120) < 1e-5
