"""Mesh pre-processing for the host model compiler.

Replaces what ``mj.MjModel.from_xml_path`` (reference call site
``envs/nightmare_v3_env.py:37``) does to every ``<mesh>`` asset of
``models/nightmare_v3/mjmodel.xml:5-23``: binary-STL load, vertex de-duplication,
scaling, mass properties ("legacy" MuJoCo-3.1.2 rule) and a
convex hull with its vertex adjacency graph (used by the plane-hull support search).

Host-side, runs once per model; numpy + scipy's qhull wrapper (the same library
MuJoCo itself links for hulls).
"""
from __future__ import annotations

import struct

import numpy as np


def load_stl_binary(path: str) -> np.ndarray:
    """Return the triangle soup of a binary STL as float32 ``[ntri, 3, 3]``."""
    with open(path, "rb") as fh:
        raw = fh.read()
    if len(raw) < 84:
        raise ValueError(f"{path}: not a binary STL (too short)")
    (ntri,) = struct.unpack_from("<I", raw, 80)
    if 84 + 50 * ntri != len(raw):
        raise ValueError(f"{path}: binary STL size mismatch ({ntri} faces, {len(raw)} bytes)")
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("attr", "<u2")])
    tris = np.frombuffer(raw, dtype=rec, count=ntri, offset=84)["v"]
    return np.ascontiguousarray(tris)


def dedup_vertices(tris: np.ndarray):
    """Merge bit-identical float32 vertices. Returns (verts float64 [nv,3], faces int32 [nf,3]).

    First-occurrence order is kept so vertex ids are reproducible.
    """
    flat = tris.reshape(-1, 3)
    keys = flat.view(np.uint32).reshape(-1, 3)
    _, first, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique-id -> rank by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    verts = flat[first[order]].astype(np.float64)
    faces = rank[inverse.reshape(-1)].reshape(-1, 3).astype(np.int32)
    return verts, faces


def _tet_moments(a, b, c, d, vol):
    """Second-moment (covariance about the origin) of tetrahedra (a,b,c,d) with given volumes."""
    s = a + b + c + d
    outer = lambda x, y: x[:, :, None] * y[:, None, :]
    acc = outer(a, a) + outer(b, b) + outer(c, c) + outer(d, d) + outer(s, s)
    return (vol[:, None, None] / 20.0 * acc).sum(axis=0)


def mass_properties(verts: np.ndarray, faces: np.ndarray, exact: bool = False):
    """Volume, centre of mass and inertia tensor (unit density, about the COM).

    ``exact=False`` is the MuJoCo 3.1.2 default (``exactmeshinertia`` unset): tetrahedra are
    spanned from the area-weighted mean of the face centroids and every tetrahedron counts with
    its ABSOLUTE volume, which over-estimates non-convex CAD meshes (SURVEY.md hard part #2).
    ``exact=True`` uses signed volumes from the origin (only correct for watertight meshes).
    """
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    nrm = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    keep = area > 1e-30
    a, b, c, nrm, area = a[keep], b[keep], c[keep], nrm[keep], area[keep]
    fcen = (a + b + c) / 3.0
    if exact:
        apex = np.zeros(3)
    else:
        apex = (fcen * area[:, None]).sum(axis=0) / area.sum()
    d = np.broadcast_to(apex, a.shape)
    vol = np.einsum("ij,ij->i", a - d, np.cross(b - d, c - d)) / 6.0
    if not exact:
        vol = np.abs(vol)
    volume = vol.sum()
    if volume <= 0:
        raise ValueError("mesh volume is not positive")
    com = ((a + b + c + d) / 4.0 * vol[:, None]).sum(axis=0) / volume
    cov = _tet_moments(a - com, b - com, c - com, d - com, vol)
    inertia = np.trace(cov) * np.eye(3) - cov
    return float(volume), com, inertia


def principal_axes(inertia: np.ndarray):
    """Diagonalise a symmetric inertia tensor -> (moments descending, right-handed rotation)."""
    w, v = np.linalg.eigh(0.5 * (inertia + inertia.T))
    idx = np.argsort(-w, kind="stable")
    w, v = w[idx], v[:, idx]
    # deterministic signs: largest-|component| of each axis positive, then fix handedness
    for k in range(2):
        j = int(np.argmax(np.abs(v[:, k])))
        if v[j, k] < 0:
            v[:, k] = -v[:, k]
    v[:, 2] = np.cross(v[:, 0], v[:, 1])
    return w, v


def convex_hull_graph(points: np.ndarray):
    """Convex hull of a point cloud -> (hull vertex ids into ``points``, CSR adjacency).

    Adjacency is expressed in LOCAL hull ids (0..nh-1), neighbours sorted ascending; it is the
    edge graph of qhull's triangulated facets (the structure MuJoCo stores as ``mesh_graph``).
    """
    from scipy.spatial import ConvexHull  # qhull

    hull = ConvexHull(points, qhull_options="Qt")
    ids = np.sort(hull.vertices).astype(np.int32)
    local = -np.ones(points.shape[0], dtype=np.int64)
    local[ids] = np.arange(ids.size)
    nbr = [set() for _ in range(ids.size)]
    for tri in hull.simplices:
        la = local[tri]
        for i in range(3):
            p, q = int(la[i]), int(la[(i + 1) % 3])
            nbr[p].add(q)
            nbr[q].add(p)
    adr = np.zeros(ids.size + 1, dtype=np.int32)
    flat = []
    for i, s in enumerate(nbr):
        lst = sorted(s)
        flat.extend(lst)
        adr[i + 1] = len(flat)
    return ids, adr, np.asarray(flat, dtype=np.int32), hull.simplices.shape[0]
