"""Host model compiler (nightmare_rl_b200/mjcf.py) — replaces MjModel.from_xml_path
(reference envs/nightmare_v3_env.py:37).  Ground truth: SURVEY.md Appendix B + scipy's qhull."""
import os

import numpy as np
import pytest

from nightmare_rl_b200 import meshproc, mjcf
from conftest import REF_MODELS

needs_ref = pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference assets not present on this box")


def test_sizes(compiled_model):
    m = compiled_model
    assert (m.nq, m.nv, m.nu, m.nbody, m.nsensor) == (25, 24, 18, 20, 13)
    assert m.nsite == 25
    # collision-capable geoms: floor, base hull, six tibia hulls (coxa/femur are contype=conaffinity=0)
    assert m.ngeom == 8
    assert list(m.geom_type) == [mjcf.GEOM_PLANE] + [mjcf.GEOM_MESH] * 7


def test_options(compiled_model):
    m = compiled_model
    assert list(m.opt_int[:5]) == [mjcf.INT_IMPLICITFAST, mjcf.SOL_PGS, mjcf.CONE_PYRAMIDAL, 3, 4]
    assert m.opt_real[0] == 0.008 and m.opt_real[3] == -9.81
    assert m.opt_real[4] == 1e-8 and m.opt_real[5] == 1e-6


def test_mass_properties(compiled_model):
    m = compiled_model
    assert abs(m.body_mass.sum() - 3.0) < 1e-12           # settotalmass=3 (mjmodel.xml:2)
    # legacy mesh-inertia mass split
    assert abs(m.body_mass[1] - 1.766) < 2e-3
    assert np.allclose(m.body_mass[[2, 3, 4]], [0.0386, 0.0472, 0.120], atol=5e-4)
    assert (m.body_inertia[1:] > 0).all()
    assert (m.body_inertia[1:, 0] >= m.body_inertia[1:, 1]).all() and (m.body_inertia[1:, 1] >= m.body_inertia[1:, 2]).all()
    q = m.body_iquat
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0)


def test_tree_and_actuators(compiled_model):
    m = compiled_model
    assert list(m.body_parent[:5]) == [-1, 0, 1, 2, 3]
    assert m.jnt_type[0] == mjcf.JNT_FREE and (m.jnt_type[1:] == mjcf.JNT_HINGE).all()
    assert list(m.act_dof) == list(range(6, 24))
    assert np.allclose(m.act_gain[:, 0], 0.8) and np.allclose(m.act_bias[:, 2], -0.8)   # velocity servo kv=0.8
    assert (m.act_ctrllimited == 1).all() and np.allclose(m.act_ctrlrange, [[-8, 8]] * 18)
    assert np.allclose(m.qpos0[:7], [0, 0, 0.15, 1, 0, 0, 0]) and np.allclose(m.qpos0[7:], 0)
    assert np.allclose(m.jnt_axis[1], [0, 0, -1])


def test_collision_filter(compiled_model):
    m = compiled_model
    # floor 1/1, base 0/1, tibia 2/3: all seven hulls collide with the floor
    assert list(m.geom_plane) == [-1, 0, 0, 0, 0, 0, 0, 0]
    assert list(m.geom_body) == [0, 1, 4, 7, 10, 13, 16, 19]


def test_sensor_layout(compiled_model):
    m = compiled_model
    names = m.names["sensor"]
    assert names[:6] == [f"leg_{k}_tibia" for k in range(1, 7)]
    assert names[6:12] == [f"leg_{k}_foot" for k in range(1, 7)] and names[12] == "base_link"
    r = m.site_size[m.sensor_site]
    assert np.allclose(r[:6], 10) and np.allclose(r[6:12], 0.007) and r[12] == 10


def test_hull_graph_is_consistent(compiled_model):
    m = compiled_model
    assert list(m.geom_hull_num[1:]) == [154, 253, 284, 252, 253, 284, 252] or m.geom_hull_num[1] == 154
    for g in range(1, m.ngeom):
        adr, num = m.geom_hull_adr[g], m.geom_hull_num[g]
        v = m.hull_vert[adr:adr + num].astype(np.float64)
        for i in range(num):
            nb = m.hull_nbr[m.hull_nbr_adr[adr + i]:m.hull_nbr_adr[adr + i + 1]]
            assert len(nb) >= 3 and (nb >= 0).all() and (nb < num).all() and i not in nb
        # convexity: hill-climbing from vertex 0 reaches the exhaustive argmin for random directions
        rng = np.random.default_rng(g)
        for _ in range(20):
            d = rng.normal(size=3)
            cur = 0
            while True:
                nb = m.hull_nbr[m.hull_nbr_adr[adr + cur]:m.hull_nbr_adr[adr + cur + 1]]
                best = nb[np.argmin(v[nb] @ d)]
                if v[best] @ d < v[cur] @ d:
                    cur = best
                else:
                    break
            assert abs(v[cur] @ d - (v @ d).min()) < 1e-9


def test_fk_at_qpos0(compiled_model, oracle_model):
    """Coordinates of SURVEY.md Appendix B: coxa/femur/tibia origins, lowest hull vertices."""
    from oracle import oracle as O
    b = O.OracleBatch(oracle_model, 1)
    b.forward(np.zeros((1, 18)))
    xpos = b.get(0, "xpos").reshape(-1, 3)
    xmat = b.get(0, "xmat").reshape(-1, 3, 3)
    assert np.allclose(xpos[1], [0, 0, 0.15])
    assert np.allclose(xpos[[2, 5, 8, 11, 14, 17], 2], 0.1955, atol=1e-6)
    assert np.allclose(xpos[[3, 6, 9, 12, 15, 18], 2], 0.167, atol=1e-6)
    assert np.allclose(xpos[[4, 7, 10, 13, 16, 19], 2], 0.2704, atol=1e-4)
    m = compiled_model
    lows = []
    for g in range(1, m.ngeom):
        bid = m.geom_body[g]
        v = m.hull_vert[m.geom_hull_adr[g]:m.geom_hull_adr[g] + m.geom_hull_num[g]].astype(np.float64)
        lows.append((xpos[bid] + v @ xmat[bid].T)[:, 2].min())
    assert abs(lows[0] - 0.1386) < 2e-4                  # base
    assert abs(min(lows[1:]) - 0.1277) < 2e-4            # tibias
    site = b.get(0, "site_xpos").reshape(-1, 3)
    feet = site[m.sensor_site[6:12]]
    assert np.allclose(feet[:, 2], 0.131, atol=1.5e-3)
    assert ((np.linalg.norm(feet[:, :2], axis=1) > 0.26) & (np.linalg.norm(feet[:, :2], axis=1) < 0.28)).all()


def test_invweight_and_meaninertia(compiled_model, oracle_model):
    """body_invweight0 / meaninertia from the compiler's own numpy mass matrix agree with the oracle's CRBA."""
    from oracle import oracle as O
    b = O.OracleBatch(oracle_model, 1)
    b.forward(np.zeros((1, 18)))
    M = b.get(0, "M").reshape(24, 24)
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    assert abs(np.trace(M) / 24 - compiled_model.opt_real[7]) < 1e-10
    assert (compiled_model.body_invweight0[1:] > 0).all()


def test_nmb_roundtrip(compiled_model):
    raw = compiled_model.to_bytes()
    again = mjcf.CompiledModel.from_bytes(raw)
    assert again.names == compiled_model.names
    for k, v in compiled_model.arrays.items():
        assert np.array_equal(v, again.arrays[k]) and v.dtype == again.arrays[k].dtype


def test_name2id(compiled_model):
    assert compiled_model.name2id(mjcf.OBJ_BODY, "base_link") == 1       # env.py:48
    assert compiled_model.name2id(mjcf.OBJ_BODY, "nope") == -1
    assert compiled_model.name2id(mjcf.OBJ_JOINT, "leg1coxa") == 1


def test_errors(tmp_path):
    with pytest.raises(mjcf.MJCFError):
        mjcf.compile_mjcf(str(tmp_path / "missing.xml"))
    bad = tmp_path / "bad.xml"
    bad.write_text("<mujoco><worldbody><body><joint type='ball'/></body></worldbody></mujoco>")
    with pytest.raises(mjcf.MJCFError):
        mjcf.compile_mjcf(str(bad))
    with pytest.raises(mjcf.MJCFError):
        mjcf.CompiledModel.from_bytes(b"XXXX0000")


@needs_ref
def test_committed_nmb_matches_fresh_compile(compiled_model):
    fresh = mjcf.compile_mjcf(os.path.join(REF_MODELS, "nightmare_v3", "mjmodel.xml"))
    for k, v in fresh.arrays.items():
        assert np.array_equal(v, compiled_model.arrays[k]), k


@needs_ref
def test_mesh_against_scipy_and_appendix_b():
    tris = meshproc.load_stl_binary(os.path.join(REF_MODELS, "nightmare_v3", "base_link.stl"))
    v, f = meshproc.dedup_vertices(tris)
    assert (len(tris), len(v)) == (22354, 10997)
    v = v * 0.001
    vol, com, inertia = meshproc.mass_properties(v, f)
    vol_exact, _, _ = meshproc.mass_properties(v, f, exact=True)
    assert abs(vol * 1e6 - 3098.6) < 0.5 and abs(vol_exact * 1e6 - 716.1) < 0.5
    ids, adr, nbr, nfaces = meshproc.convex_hull_graph(v)
    assert (len(ids), nfaces) == (154, 304)
    from scipy.spatial import ConvexHull
    assert set(ids.tolist()) == set(ConvexHull(v).vertices.tolist())


@needs_ref
def test_other_reference_models_parse():
    mjx = mjcf.compile_mjcf(os.path.join(REF_MODELS, "nightmare_v3", "mjmodel_mjx.xml"))
    assert (mjx.nq, mjx.nv, mjx.nsensor) == (25, 24, 6)
    assert mjx.opt_int[1] == mjcf.SOL_NEWTON and mjx.opt_real[0] == 0.001 and mjx.opt_int[5] == 0
    any_ = mjcf.compile_mjcf(os.path.join(REF_MODELS, "anymal_c", "scene.xml"))
    assert (any_.nbody, any_.nv, any_.nu) == (14, 18, 12)
    assert any_.opt_int[2] == mjcf.CONE_ELLIPTIC and any_.opt_real[6] == 100
    assert np.allclose(any_.act_gain[:, 0], 100) and np.allclose(any_.act_forcerange, [[-80, 80]] * 12)
    assert np.allclose(any_.dof_damping[6:], 1.0) and np.allclose(any_.dof_frictionloss[6:], 0.1)
