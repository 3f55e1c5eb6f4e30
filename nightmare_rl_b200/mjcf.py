"""Host model compiler: MJCF (+ meshes) -> flat constant arrays ("NMB" compiled model).

This replaces ``mj.MjModel.from_xml_path(cfg.env.model_path)`` (reference
``envs/nightmare_v3_env.py:37``) for the models shipped by the reference
(``models/nightmare_v3/mjmodel.xml``, ``models/anymal_c/scene.xml``).  It is a from-scratch
compiler, not a MuJoCo port: it understands the MJCF subset those files use and emits the flat
arrays that both the CUDA step kernels (``csrc/``) and the CPU oracle (``oracle/``) consume.

Semantics follow the MuJoCo 3.1.2 documentation as summarised in SURVEY.md Appendix A.1:
radian/degree angles, default classes, ``inertiafromgeom`` + ``settotalmass``, legacy mesh
inertia, convex hulls for colliding meshes, ``body_invweight0`` / ``meaninertia`` at ``qpos0``.

The compiled model can be stored as an ``.nmb`` file (the analogue of MuJoCo's ``.mjb``):
a list of named little-endian arrays, readable from C without any dependency.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from . import meshproc

# ----------------------------------------------------------------------------- constants
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 0, 2, 3, 5, 6, 7
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
               "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH}
INT_EULER, INT_RK4, INT_IMPLICIT, INT_IMPLICITFAST = 0, 1, 2, 3
SOL_PGS, SOL_CG, SOL_NEWTON = 0, 1, 2
CONE_PYRAMIDAL, CONE_ELLIPTIC = 0, 1
OBJ_BODY, OBJ_JOINT, OBJ_GEOM, OBJ_SITE, OBJ_ACTUATOR, OBJ_SENSOR = 1, 3, 5, 6, 19, 20  # mjtObj ids
MINVAL = 1e-15

NMB_MAGIC = b"NMB1"


# ----------------------------------------------------------------------------- quaternion helpers
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def mat_to_quat(m):
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(m)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
        q = np.zeros(4)
        q[0] = (m[k, j] - m[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (m[j, i] + m[i, j]) / s
        q[1 + k] = (m[k, i] + m[i, k]) / s
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def _unit_quat(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n < MINVAL:
        raise ValueError("zero quaternion in MJCF")
    return q / n


# ----------------------------------------------------------------------------- spec records
@dataclass
class _Body:
    name: str
    parent: int
    pos: np.ndarray
    quat: np.ndarray
    joints: list = field(default_factory=list)
    geoms: list = field(default_factory=list)
    sites: list = field(default_factory=list)
    inertial: dict | None = None


class MJCFError(ValueError):
    """Raised for malformed or unsupported MJCF (the reference raises mujoco.FatalError)."""


# ----------------------------------------------------------------------------- parser
class _Parser:
    def __init__(self, path: str):
        self.path = os.path.abspath(path)
        self.dir = os.path.dirname(self.path)
        if not os.path.exists(self.path):
            raise MJCFError(f"MJCF file not found: {path}")
        self.root = self._load(self.path)
        self.degree = True
        self.meshdir = ""
        self.settotalmass = -1.0
        self.inertiafromgeom = "auto"
        self.autolimits = False
        self.exactmeshinertia = False
        self.defaults: dict[str, dict[str, dict[str, str]]] = {"main": {}}
        self.default_parent: dict[str, str | None] = {"main": None}
        self.meshes: dict[str, dict] = {}
        self.bodies: list[_Body] = []
        self.actuators: list[dict] = []
        self.sensors: list[dict] = []
        self.excludes: list[tuple[str, str]] = []
        self.option = dict(timestep=0.002, gravity=[0, 0, -9.81], integrator="Euler", iterations=100,
                           ls_iterations=50, noslip_iterations=0, tolerance=1e-8, noslip_tolerance=1e-6,
                           solver="Newton", cone="pyramidal", impratio=1.0, eulerdamp=True)

    # -- xml loading with <include>
    def _load(self, path):
        try:
            root = ET.parse(path).getroot()
        except ET.ParseError as exc:
            raise MJCFError(f"XML parse error in {path}: {exc}") from exc
        if root.tag != "mujoco":
            raise MJCFError(f"{path}: root element must be <mujoco>")
        self._expand_includes(root, os.path.dirname(path))
        return root

    def _expand_includes(self, node, base):
        i = 0
        while i < len(node):
            ch = node[i]
            if ch.tag == "include":
                inc = ET.parse(os.path.join(base, ch.attrib["file"])).getroot()
                self._expand_includes(inc, base)
                node.remove(ch)
                for k, sub in enumerate(list(inc)):
                    node.insert(i + k, sub)
                i += len(list(inc))
            else:
                self._expand_includes(ch, base)
                i += 1

    # -- attribute helpers
    @staticmethod
    def _floats(s, n=None):
        v = np.array([float(x) for x in s.split()], dtype=np.float64)
        if n is not None and v.size != n:
            raise MJCFError(f"expected {n} numbers, got '{s}'")
        return v

    def _resolve(self, tag, elem, klass):
        """Merge default-class attributes (outermost first) with the element's own."""
        chain = []
        c = elem.attrib.get("class", klass) or "main"
        if c not in self.defaults:
            raise MJCFError(f"unknown default class '{c}'")
        while c is not None:
            chain.append(c)
            c = self.default_parent[c]
        out: dict[str, str] = {}
        for c in reversed(chain):
            out.update(self.defaults[c].get(tag, {}))
        out.update(elem.attrib)
        return out

    def _parse_defaults(self, node, name, parent):
        if name not in self.defaults:
            self.defaults[name] = {}
            self.default_parent[name] = parent
        for ch in node:
            if ch.tag == "default":
                self._parse_defaults(ch, ch.attrib["class"], name)
            else:
                tag = ch.tag
                self.defaults[name].setdefault(tag, {}).update(ch.attrib)
                if tag in ("position", "velocity", "motor", "general"):
                    self.defaults[name].setdefault("actuator_any", {}).update(ch.attrib)

    def _orientation(self, a):
        if "quat" in a:
            return _unit_quat(self._floats(a["quat"], 4))
        if "euler" in a:
            e = self._floats(a["euler"], 3)
            if self.degree:
                e = np.deg2rad(e)
            q = np.array([1.0, 0, 0, 0])
            for ax, ang in zip(np.eye(3), e):  # default eulerseq "xyz", intrinsic
                q = quat_mul(q, np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax]))
            return q
        if "axisangle" in a:
            v = self._floats(a["axisangle"], 4)
            ang = np.deg2rad(v[3]) if self.degree else v[3]
            ax = v[:3] / np.linalg.norm(v[:3])
            return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        return np.array([1.0, 0, 0, 0])

    # -- sections
    def parse(self):
        r = self.root
        for comp in r.findall("compiler"):
            a = comp.attrib
            if "angle" in a:
                self.degree = a["angle"] == "degree"
            self.meshdir = a.get("meshdir", self.meshdir)
            if "settotalmass" in a:
                self.settotalmass = float(a["settotalmass"])
            self.inertiafromgeom = a.get("inertiafromgeom", self.inertiafromgeom)
            self.autolimits = a.get("autolimits", "false") == "true" or self.autolimits
            self.exactmeshinertia = a.get("exactmeshinertia", "false") == "true"
        for opt in r.findall("option"):
            a = opt.attrib
            for k in ("timestep", "tolerance", "noslip_tolerance", "impratio"):
                if k in a:
                    self.option[k] = float(a[k])
            for k in ("iterations", "ls_iterations", "noslip_iterations"):
                if k in a:
                    self.option[k] = int(a[k])
            for k in ("integrator", "solver", "cone"):
                if k in a:
                    self.option[k] = a[k]
            if "gravity" in a:
                self.option["gravity"] = list(self._floats(a["gravity"], 3))
            for fl in opt.findall("flag"):
                if fl.attrib.get("eulerdamp") == "disable":
                    self.option["eulerdamp"] = False
        for d in r.findall("default"):
            self._parse_defaults(d, d.attrib.get("class", "main"), None if "class" not in d.attrib else "main")
        for asset in r.findall("asset"):
            for m in asset.findall("mesh"):
                a = self._resolve("mesh", m, "main")
                fname = a.get("file")
                name = a.get("name") or os.path.splitext(os.path.basename(fname))[0]
                self.meshes[name] = dict(file=os.path.join(self.dir, self.meshdir, fname),
                                         scale=self._floats(a.get("scale", "1 1 1"), 3))
        wb = r.findall("worldbody")
        self.bodies.append(_Body("world", -1, np.zeros(3), np.array([1.0, 0, 0, 0])))
        for w in wb:
            self._parse_body_children(w, 0, "main")
        for act in r.findall("actuator"):
            for e in act:
                self.actuators.append(self._parse_actuator(e))
        for sen in r.findall("sensor"):
            for e in sen:
                self.sensors.append(dict(type=e.tag, **e.attrib))
        for con in r.findall("contact"):
            for e in con.findall("exclude"):
                self.excludes.append((e.attrib["body1"], e.attrib["body2"]))
        return self

    def _parse_body_children(self, node, bid, klass):
        for ch in node:
            if ch.tag == "body":
                k2 = ch.attrib.get("childclass", klass)
                b = _Body(ch.attrib.get("name", f"body{len(self.bodies)}"), bid,
                          self._floats(ch.attrib.get("pos", "0 0 0"), 3), self._orientation(ch.attrib))
                self.bodies.append(b)
                self._parse_body_children(ch, len(self.bodies) - 1, k2)
            elif ch.tag in ("joint", "freejoint"):
                a = self._resolve("joint", ch, klass) if ch.tag == "joint" else dict(ch.attrib, type="free")
                self.bodies[bid].joints.append(a)
            elif ch.tag == "geom":
                self.bodies[bid].geoms.append(self._resolve("geom", ch, klass))
            elif ch.tag == "site":
                self.bodies[bid].sites.append(self._resolve("site", ch, klass))
            elif ch.tag == "inertial":
                self.bodies[bid].inertial = dict(ch.attrib)

    def _parse_actuator(self, e):
        klass = e.attrib.get("class", "main")
        a = self._resolve(e.tag, e, klass)
        gain, bias = np.zeros(3), np.zeros(3)
        if e.tag == "motor":
            gain[0] = 1.0
        elif e.tag == "position":
            kp = float(a.get("kp", 1.0))
            kv = float(a.get("kv", 0.0))
            gain[0] = kp
            bias[:] = [0, -kp, -kv]
        elif e.tag == "velocity":
            kv = float(a.get("kv", 1.0))
            gain[0] = kv
            bias[:] = [0, 0, -kv]
        elif e.tag == "general":
            if "gainprm" in a:
                g = self._floats(a["gainprm"])
                gain[:min(3, g.size)] = g[:3]
            else:
                gain[0] = 1.0
            if "biasprm" in a:
                b = self._floats(a["biasprm"])
                bias[:min(3, b.size)] = b[:3]
        else:
            raise MJCFError(f"unsupported actuator <{e.tag}>")
        out = dict(name=a.get("name", ""), joint=a.get("joint"), gain=gain, bias=bias,
                   gear=self._floats(a.get("gear", "1"))[0])
        for key in ("ctrl", "force"):
            rng = self._floats(a[key + "range"], 2) if key + "range" in a else np.zeros(2)
            lim = a.get(key + "limited", "auto")
            limited = (lim == "true") or (lim == "auto" and self.autolimits and key + "range" in a)
            out[key + "range"], out[key + "limited"] = rng, bool(limited)
        if out["joint"] is None:
            raise MJCFError("only joint transmissions are supported")
        return out


# ----------------------------------------------------------------------------- compiled model container
class CompiledModel:
    """Flat, named numpy arrays + a few python-side name tables. See ``FIELDS`` in the .nmb."""

    def __init__(self, arrays: dict[str, np.ndarray], names: dict[str, list[str]]):
        self.arrays = arrays
        self.names = names

    def __getattr__(self, k):
        try:
            return self.__dict__["arrays"][k]
        except KeyError as exc:
            raise AttributeError(k) from exc

    # sizes as python ints
    @property
    def nq(self): return int(self.arrays["sizes"][0])
    @property
    def nv(self): return int(self.arrays["sizes"][1])
    @property
    def nu(self): return int(self.arrays["sizes"][2])
    @property
    def nbody(self): return int(self.arrays["sizes"][3])
    @property
    def ngeom(self): return int(self.arrays["sizes"][5])
    @property
    def nsite(self): return int(self.arrays["sizes"][6])
    @property
    def nsensor(self): return int(self.arrays["sizes"][7])

    def name2id(self, objtype: int, name: str) -> int:
        """≙ ``mj.mj_name2id`` (reference ``envs/nightmare_v3_env.py:48``); -1 when absent."""
        key = {OBJ_BODY: "body", OBJ_JOINT: "joint", OBJ_GEOM: "geom", OBJ_SITE: "site",
               OBJ_ACTUATOR: "actuator", OBJ_SENSOR: "sensor"}.get(objtype)
        if key is None:
            return -1
        try:
            return self.names[key].index(name)
        except ValueError:
            return -1

    # ---- .nmb serialisation: magic, count, then records (name[32], dtype code, ndim, dims[4], nbytes, data, pad to 8)
    _CODES = {np.dtype("<f8"): 0, np.dtype("<f4"): 1, np.dtype("<i4"): 2, np.dtype("u1"): 3}

    def to_bytes(self) -> bytes:
        arrays = dict(self.arrays)
        for key, lst in self.names.items():
            arrays["names_" + key] = np.frombuffer(("\n".join(lst)).encode(), dtype=np.uint8)
        out = [NMB_MAGIC, struct.pack("<I", len(arrays))]
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            code = self._CODES[arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype]
            dims = list(arr.shape) + [0] * (4 - arr.ndim)
            nm = name.encode()
            if len(nm) > 31:
                raise ValueError(name)
            out.append(struct.pack("<32sII4qq", nm, code, arr.ndim, *dims, arr.nbytes))
            out.append(arr.tobytes())
            out.append(b"\0" * ((-arr.nbytes) % 8))
        return b"".join(out)

    def save(self, path: str) -> None:
        with open(path, "wb") as fh:
            fh.write(self.to_bytes())

    @classmethod
    def from_bytes(cls, raw: bytes) -> "CompiledModel":
        if raw[:4] != NMB_MAGIC:
            raise MJCFError("not an NMB compiled model")
        (cnt,) = struct.unpack_from("<I", raw, 4)
        off = 8
        inv = {v: k for k, v in cls._CODES.items()}
        arrays, names = {}, {}
        for _ in range(cnt):
            nm, code, ndim, d0, d1, d2, d3, nbytes = struct.unpack_from("<32sII4qq", raw, off)
            off += struct.calcsize("<32sII4qq")
            shape = (d0, d1, d2, d3)[:ndim]
            arr = np.frombuffer(raw, dtype=inv[code], count=nbytes // inv[code].itemsize, offset=off).reshape(shape).copy()
            off += nbytes + ((-nbytes) % 8)
            name = nm.rstrip(b"\0").decode()
            if name.startswith("names_"):
                s = arr.tobytes().decode()
                names[name[6:]] = s.split("\n") if s else []
            else:
                arrays[name] = arr
        return cls(arrays, names)

    @classmethod
    def load(cls, path: str) -> "CompiledModel":
        with open(path, "rb") as fh:
            return cls.from_bytes(fh.read())


# ----------------------------------------------------------------------------- primitive mass properties
def _primitive_mass(gtype, size, density):
    if gtype == GEOM_SPHERE:
        r = size[0]
        m = density * 4 / 3 * np.pi * r ** 3
        return m, np.full(3, 0.4 * m * r * r)
    if gtype == GEOM_BOX:
        m = density * 8 * size[0] * size[1] * size[2]
        return m, m / 3 * np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2, size[0] ** 2 + size[1] ** 2])
    if gtype == GEOM_CYLINDER:
        r, h = size[0], size[1]
        m = density * np.pi * r * r * 2 * h
        ix = m * (3 * r * r + (2 * h) ** 2) / 12
        return m, np.array([ix, ix, 0.5 * m * r * r])
    if gtype == GEOM_CAPSULE:
        r, h = size[0], size[1]
        mc = density * np.pi * r * r * 2 * h
        ms = density * 4 / 3 * np.pi * r ** 3
        ix = mc * (3 * r * r + 4 * h * h) / 12 + ms * (0.4 * r * r + h * h + 0.75 * r * h)
        return mc + ms, np.array([ix, ix, 0.5 * mc * r * r + 0.4 * ms * r * r])
    return 0.0, np.zeros(3)


# ----------------------------------------------------------------------------- compile
def compile_mjcf(path: str) -> CompiledModel:
    """Compile an MJCF file into a :class:`CompiledModel` (≙ ``MjModel.from_xml_path``)."""
    p = _Parser(path).parse()
    F = p._floats
    nbody = len(p.bodies)
    body_names = [b.name for b in p.bodies]

    # ---------------- joints / dofs
    jnt_type, jnt_body, jnt_qposadr, jnt_dofadr, jnt_pos, jnt_axis = [], [], [], [], [], []
    jnt_limited, jnt_range, jnt_names = [], [], []
    dof_body, dof_jnt, dof_parent, dof_damping, dof_frictionloss, dof_armature = [], [], [], [], [], []
    body_jntadr, body_jntnum, body_dofadr, body_dofnum = [], [], [], []
    qpos0 = []
    body_lastdof = [-1] * nbody
    for bid, b in enumerate(p.bodies):
        body_jntadr.append(len(jnt_type) if b.joints else -1)
        body_jntnum.append(len(b.joints))
        body_dofadr.append(len(dof_body) if b.joints else -1)
        last = body_lastdof[b.parent] if b.parent >= 0 else -1
        nd0 = len(dof_body)
        for a in b.joints:
            t = a.get("type", "hinge")
            jid = len(jnt_type)
            jnt_names.append(a.get("name", ""))
            jnt_body.append(bid)
            jnt_qposadr.append(len(qpos0))
            jnt_dofadr.append(len(dof_body))
            jnt_pos.append(F(a.get("pos", "0 0 0"), 3))
            ax = F(a.get("axis", "0 0 1"), 3)
            jnt_axis.append(ax / max(np.linalg.norm(ax), MINVAL))
            rng = F(a["range"], 2) if "range" in a else np.zeros(2)
            if "range" in a and p.degree and t == "hinge":
                rng = np.deg2rad(rng)
            lim = a.get("limited", "auto")
            jnt_limited.append(int(lim == "true" or (lim == "auto" and p.autolimits and "range" in a)))
            jnt_range.append(rng)
            damping, floss, arm = float(a.get("damping", 0)), float(a.get("frictionloss", 0)), float(a.get("armature", 0))
            if t == "free":
                jnt_type.append(JNT_FREE)
                qpos0.extend(list(b.pos) + list(b.quat))
                nd = 6
            elif t == "hinge":
                jnt_type.append(JNT_HINGE)
                qpos0.append(float(a.get("ref", 0)))
                nd = 1
            elif t == "slide":
                jnt_type.append(JNT_SLIDE)
                qpos0.append(float(a.get("ref", 0)))
                nd = 1
            else:
                raise MJCFError(f"unsupported joint type '{t}'")
            for _ in range(nd):
                dof_body.append(bid)
                dof_jnt.append(jid)
                dof_parent.append(last)
                last = len(dof_body) - 1
                dof_damping.append(damping)
                dof_frictionloss.append(floss)
                dof_armature.append(arm)
        body_dofnum.append(len(dof_body) - nd0)
        body_lastdof[bid] = last
    nq, nv, njnt = len(qpos0), len(dof_body), len(jnt_type)
    qpos0 = np.array(qpos0, dtype=np.float64)

    body_rootid = list(range(nbody))
    for bid in range(1, nbody):
        par = p.bodies[bid].parent
        body_rootid[bid] = bid if par == 0 else body_rootid[par]

    # ---------------- geoms (mass + collision)
    mesh_cache: dict[str, dict] = {}

    def mesh_data(name):
        if name not in p.meshes:
            raise MJCFError(f"unknown mesh '{name}'")
        if name not in mesh_cache:
            spec = p.meshes[name]
            ext = os.path.splitext(spec["file"])[1].lower()
            if ext != ".stl":
                raise MJCFError(f"mesh '{name}': only binary STL is supported for mass/collision geoms ({ext})")
            v, f = meshproc.dedup_vertices(meshproc.load_stl_binary(spec["file"]))
            v = v * spec["scale"]
            vol, com, inertia = meshproc.mass_properties(v, f, exact=p.exactmeshinertia)
            mesh_cache[name] = dict(verts=v, faces=f, volume=vol, com=com, inertia=inertia)
        return mesh_cache[name]

    geoms = []          # collision-capable geoms only
    geom_names = []
    hull_vert, hull_nbr_adr, hull_nbr = [], [0], []
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    body_inertia = np.zeros((nbody, 3))

    for bid, b in enumerate(p.bodies):
        use_geoms = p.inertiafromgeom == "true" or (p.inertiafromgeom == "auto" and b.inertial is None)
        acc_m, acc_mc, parts = 0.0, np.zeros(3), []
        for a in b.geoms:
            gtype = _GEOM_TYPES.get(a.get("type", "sphere"))
            if gtype is None:
                raise MJCFError(f"unsupported geom type '{a.get('type')}'")
            size = np.zeros(3)
            if "size" in a:
                s = F(a["size"])
                size[:s.size] = s[:3]
            gpos = F(a.get("pos", "0 0 0"), 3)
            gquat = p._orientation(a)
            if "fromto" in a:
                ft = F(a["fromto"], 6)
                gpos = 0.5 * (ft[:3] + ft[3:])
                d = ft[3:] - ft[:3]
                size[1] = 0.5 * np.linalg.norm(d)
                z = d / np.linalg.norm(d)
                v = np.cross([0, 0, 1.0], z)
                s_, c_ = np.linalg.norm(v), z[2]
                gquat = np.array([1.0, 0, 0, 0]) if s_ < 1e-12 else np.concatenate(
                    [[np.cos(np.arctan2(s_, c_) / 2)], np.sin(np.arctan2(s_, c_) / 2) * v / s_])
            gmat = quat_to_mat(gquat)
            contype, conaff = int(a.get("contype", 1)), int(a.get("conaffinity", 1))
            density = float(a.get("density", 1000.0))
            md = mesh_data(a["mesh"]) if gtype == GEOM_MESH and (use_geoms and bid > 0 or contype or conaff) else None
            # --- mass contribution
            if use_geoms and bid > 0 and gtype != GEOM_PLANE:
                if gtype == GEOM_MESH:
                    m = density * md["volume"]
                    com_b = gpos + gmat @ md["com"]
                    inertia_b = gmat @ (density * md["inertia"]) @ gmat.T
                else:
                    m, diag = _primitive_mass(gtype, size, density)
                    com_b, inertia_b = gpos, gmat @ np.diag(diag) @ gmat.T
                if "mass" in a:
                    sc = float(a["mass"]) / m
                    m, inertia_b = m * sc, inertia_b * sc
                acc_m += m
                acc_mc += m * com_b
                parts.append((m, com_b, inertia_b))
            # --- collision record
            if contype or conaff:
                fr = np.array([1.0, 0.005, 0.0001])
                if "friction" in a:
                    fv = F(a["friction"])
                    fr[:fv.size] = fv
                solref = np.array([0.02, 1.0])
                if "solref" in a:
                    sv = F(a["solref"])
                    solref[:sv.size] = sv
                solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
                if "solimp" in a:
                    sv = F(a["solimp"])
                    solimp[:sv.size] = sv
                rec = dict(type=gtype, body=bid, pos=gpos, quat=gquat, size=size, contype=contype, conaffinity=conaff,
                           condim=int(a.get("condim", 3)), priority=int(a.get("priority", 0)), friction=fr,
                           solref=solref, solimp=solimp, margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
                           hull_adr=-1, hull_num=0, rbound=0.0)
                if gtype == GEOM_MESH:
                    ids, adr, nbr, _ = meshproc.convex_hull_graph(md["verts"])
                    hv = md["verts"][ids]
                    # rbound: distance of the farthest hull vertex from the mesh centre of mass
                    rec["rbound"] = float(np.linalg.norm(hv - md["com"], axis=1).max())
                    rec["center"] = gpos + gmat @ md["com"]
                    rec["hull_adr"], rec["hull_num"] = len(hull_vert), len(ids)
                    base_adr = hull_nbr_adr[-1]
                    hull_vert.extend((gpos + hv @ gmat.T).tolist())      # hull vertices in the BODY frame
                    hull_nbr_adr.extend((base_adr + adr[1:]).tolist())
                    hull_nbr.extend(nbr.tolist())
                elif gtype == GEOM_SPHERE:
                    rec["rbound"], rec["center"] = size[0], gpos
                elif gtype == GEOM_BOX:
                    rec["rbound"], rec["center"] = float(np.linalg.norm(size)), gpos
                elif gtype in (GEOM_CYLINDER,):
                    rec["rbound"], rec["center"] = float(np.hypot(size[0], size[1])), gpos
                elif gtype == GEOM_CAPSULE:
                    rec["rbound"], rec["center"] = float(size[0] + size[1]), gpos
                else:
                    rec["rbound"], rec["center"] = 0.0, gpos
                geoms.append(rec)
                geom_names.append(a.get("name", ""))
        if bid == 0:
            continue
        if use_geoms:
            if acc_m <= 0:
                continue
            com = acc_mc / acc_m
            tot = np.zeros((3, 3))
            for m, c, ib in parts:
                d = c - com
                tot += ib + m * (d @ d * np.eye(3) - np.outer(d, d))
            w, R = meshproc.principal_axes(tot)
            body_mass[bid], body_ipos[bid], body_inertia[bid], body_iquat[bid] = acc_m, com, w, mat_to_quat(R)
        elif b.inertial is not None:
            a = b.inertial
            body_mass[bid] = float(a["mass"])
            body_ipos[bid] = F(a.get("pos", "0 0 0"), 3)
            iq = p._orientation(a)
            if "fullinertia" in a:
                fi = F(a["fullinertia"], 6)
                full = np.array([[fi[0], fi[3], fi[4]], [fi[3], fi[1], fi[5]], [fi[4], fi[5], fi[2]]])
                w, R = meshproc.principal_axes(full)
                body_inertia[bid], iq = w, quat_mul(iq, mat_to_quat(R))
            else:
                body_inertia[bid] = F(a["diaginertia"], 3)
            body_iquat[bid] = iq
    if p.settotalmass > 0:
        sc = p.settotalmass / body_mass.sum()
        body_mass *= sc
        body_inertia *= sc

    # ---------------- sites, sensors, actuators
    site_body, site_pos, site_size, site_names = [], [], [], []
    for bid, b in enumerate(p.bodies):
        for a in b.sites:
            if a.get("type", "sphere") != "sphere":
                raise MJCFError("only sphere sites are supported")
            site_names.append(a.get("name", ""))
            site_body.append(bid)
            site_pos.append(F(a.get("pos", "0 0 0"), 3))
            site_size.append(F(a.get("size", "0.005"))[0])
    sensor_site, sensor_names = [], []
    for s in p.sensors:
        if s["type"] != "touch":
            continue  # only touch sensors are on the hot path (reference envs/nightmare_v3_env.py:224-226)
        if s["site"] not in site_names:
            raise MJCFError(f"touch sensor references unknown site '{s['site']}'")
        sensor_names.append(s.get("name", ""))
        sensor_site.append(site_names.index(s["site"]))
    act_dof, act_gain, act_bias, act_ctrlrange, act_ctrllimited, act_forcerange, act_forcelimited, act_gear = [], [], [], [], [], [], [], []
    act_names = []
    for a in p.actuators:
        if a["joint"] not in jnt_names:
            raise MJCFError(f"actuator references unknown joint '{a['joint']}'")
        jid = jnt_names.index(a["joint"])
        if jnt_type[jid] not in (JNT_HINGE, JNT_SLIDE):
            raise MJCFError("actuators must drive hinge/slide joints")
        act_names.append(a["name"])
        act_dof.append(jnt_dofadr[jid])
        act_gain.append(a["gain"]); act_bias.append(a["bias"]); act_gear.append(a["gear"])
        act_ctrlrange.append(a["ctrlrange"]); act_ctrllimited.append(int(a["ctrllimited"]))
        act_forcerange.append(a["forcerange"]); act_forcelimited.append(int(a["forcelimited"]))
    nu = len(act_dof)

    # ---------------- collision pair filter for geom-vs-plane (the pairs the step kernels handle)
    ngeom = len(geoms)
    excl = {(body_names.index(a), body_names.index(b)) for a, b in p.excludes if a in body_names and b in body_names}
    plane_ids = [i for i, g in enumerate(geoms) if g["type"] == GEOM_PLANE]
    geom_plane = -np.ones(ngeom, dtype=np.int32)     # id of the plane this geom collides with (or -1)
    for i, g in enumerate(geoms):
        if g["type"] == GEOM_PLANE:
            continue
        for pid in plane_ids:
            pg = geoms[pid]
            mask_ok = (g["contype"] & pg["conaffinity"]) or (pg["contype"] & g["conaffinity"])
            b1, b2 = pg["body"], g["body"]
            if mask_ok and b1 != b2 and (b1, b2) not in excl and (b2, b1) not in excl:
                geom_plane[i] = pid
                break

    # ---------------- kinematics + mass matrix at qpos0 -> invweight0, meaninertia
    xpos = np.zeros((nbody, 3)); xmat = np.tile(np.eye(3), (nbody, 1, 1)); xquat = np.tile([1.0, 0, 0, 0], (nbody, 1))
    for bid in range(1, nbody):
        b = p.bodies[bid]
        if any(jnt_type[j] == JNT_FREE for j in range(body_jntadr[bid], body_jntadr[bid] + body_jntnum[bid])) if b.joints else False:
            xpos[bid], xquat[bid] = b.pos, b.quat
        else:
            xpos[bid] = xpos[b.parent] + xmat[b.parent] @ b.pos
            xquat[bid] = quat_mul(xquat[b.parent], b.quat)
        xmat[bid] = quat_to_mat(xquat[bid])
    xipos = np.array([xpos[i] + xmat[i] @ body_ipos[i] for i in range(nbody)])
    ximat = np.array([xmat[i] @ quat_to_mat(body_iquat[i]) for i in range(nbody)])

    def body_jac(bid, point):
        jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
        d = body_dofadr[bid] + body_dofnum[bid] - 1 if body_dofnum[bid] else body_lastdof[bid]
        while d >= 0:
            jid = dof_jnt[d]
            jb = jnt_body[jid]
            k = d - jnt_dofadr[jid]
            if jnt_type[jid] == JNT_FREE:
                if k < 3:
                    jp[k, d] = 1.0
                else:
                    ax = xmat[jb][:, k - 3]
                    jr[:, d] = ax
                    jp[:, d] = np.cross(ax, point - xpos[jb])
            else:
                ax = xmat[jb] @ jnt_axis[jid]
                anchor = xpos[jb] + xmat[jb] @ jnt_pos[jid]
                if jnt_type[jid] == JNT_HINGE:
                    jr[:, d] = ax
                    jp[:, d] = np.cross(ax, point - anchor)
                else:
                    jp[:, d] = ax
            d = dof_parent[d]
        return jp, jr

    M = np.diag(np.array(dof_armature, dtype=np.float64)) if nv else np.zeros((0, 0))
    for bid in range(1, nbody):
        if body_mass[bid] <= 0:
            continue
        jp, jr = body_jac(bid, xipos[bid])
        Iw = ximat[bid] @ np.diag(body_inertia[bid]) @ ximat[bid].T
        M += body_mass[bid] * jp.T @ jp + jr.T @ Iw @ jr
    body_invweight0 = np.zeros((nbody, 2))
    if nv:
        Minv = np.linalg.inv(M)
        for bid in range(1, nbody):
            if body_lastdof[bid] < 0 and body_dofnum[bid] == 0 and body_rootid[bid] == bid and not p.bodies[bid].joints:
                continue
            jp, jr = body_jac(bid, xipos[bid])
            A = np.vstack([jp, jr]) @ Minv @ np.vstack([jp, jr]).T
            body_invweight0[bid] = [max(MINVAL, np.trace(A[:3, :3]) / 3), max(MINVAL, np.trace(A[3:, 3:]) / 3)]
        dof_invweight0 = np.diag(Minv).copy()
        for jid in range(njnt):
            if jnt_type[jid] == JNT_FREE:
                a0 = jnt_dofadr[jid]
                dof_invweight0[a0:a0 + 3] = dof_invweight0[a0:a0 + 3].mean()
                dof_invweight0[a0 + 3:a0 + 6] = dof_invweight0[a0 + 3:a0 + 6].mean()
        meaninertia = float(np.trace(M) / nv)
    else:
        dof_invweight0, meaninertia = np.zeros(0), 1.0

    o = p.option
    try:
        integrator = {"Euler": INT_EULER, "RK4": INT_RK4, "implicit": INT_IMPLICIT, "implicitfast": INT_IMPLICITFAST}[o["integrator"]]
        solver = {"PGS": SOL_PGS, "CG": SOL_CG, "Newton": SOL_NEWTON}[o["solver"]]
        cone = {"pyramidal": CONE_PYRAMIDAL, "elliptic": CONE_ELLIPTIC}[o["cone"]]
    except KeyError as exc:
        raise MJCFError(f"unknown option value {exc}") from exc

    i32 = lambda x, shape=None: np.asarray(x, dtype=np.int32).reshape(shape if shape is not None else -1)
    f64 = lambda x, shape: np.asarray(x, dtype=np.float64).reshape(shape)
    G = lambda key, shape, dt=np.float64: np.asarray([g[key] for g in geoms], dtype=dt).reshape(shape)
    arrays = {
        "sizes": i32([nq, nv, nu, nbody, njnt, ngeom, len(site_body), len(sensor_site), len(hull_vert), len(hull_nbr)]),
        # last entry: how many contacts one plane-mesh pair may produce (support vertex + neighbours).  4 follows SURVEY.md
        # Appendix A.2 as recalled ("up to 3 more"); the entry exists so that a MuJoCo cross-check can correct it to whatever
        # mjc_PlaneConvex really does by re-saving the model, without touching oracle or kernel code
        "opt_int": i32([integrator, solver, cone, o["iterations"], o["noslip_iterations"], int(o["eulerdamp"]), o["ls_iterations"],
                        PLANEMESH_MAXCON, MPR_ITERATIONS, PLANEMESH_ALLVERTS, PLANEMESH_SEPVERT, WARM_AFTER_NOSLIP]),
        "opt_real": f64([o["timestep"], *o["gravity"], o["tolerance"], o["noslip_tolerance"], o["impratio"], meaninertia,
                         MPR_TOLERANCE, PLANEMESH_SEP, PYRAMID_RFAC], -1),
        "qpos0": qpos0,
        "body_parent": i32([b.parent for b in p.bodies]),
        "body_rootid": i32(body_rootid),
        "body_jntadr": i32(body_jntadr), "body_jntnum": i32(body_jntnum),
        "body_dofadr": i32(body_dofadr), "body_dofnum": i32(body_dofnum),
        "body_pos": f64([b.pos for b in p.bodies], (nbody, 3)),
        "body_quat": f64([b.quat for b in p.bodies], (nbody, 4)),
        "body_ipos": body_ipos, "body_iquat": body_iquat, "body_mass": body_mass, "body_inertia": body_inertia,
        "body_invweight0": body_invweight0,
        "jnt_type": i32(jnt_type), "jnt_body": i32(jnt_body), "jnt_qposadr": i32(jnt_qposadr), "jnt_dofadr": i32(jnt_dofadr),
        "jnt_pos": f64(jnt_pos, (njnt, 3)), "jnt_axis": f64(jnt_axis, (njnt, 3)),
        "jnt_limited": i32(jnt_limited), "jnt_range": f64(jnt_range, (njnt, 2)),
        "dof_body": i32(dof_body), "dof_jnt": i32(dof_jnt), "dof_parent": i32(dof_parent),
        "dof_damping": f64(dof_damping, -1), "dof_frictionloss": f64(dof_frictionloss, -1), "dof_armature": f64(dof_armature, -1),
        "dof_invweight0": f64(dof_invweight0, -1),
        "act_dof": i32(act_dof), "act_gain": f64(act_gain, (nu, 3)), "act_bias": f64(act_bias, (nu, 3)), "act_gear": f64(act_gear, -1),
        "act_ctrlrange": f64(act_ctrlrange, (nu, 2)), "act_ctrllimited": i32(act_ctrllimited),
        "act_forcerange": f64(act_forcerange, (nu, 2)), "act_forcelimited": i32(act_forcelimited),
        "geom_type": G("type", -1, np.int32), "geom_body": G("body", -1, np.int32),
        "geom_contype": G("contype", -1, np.int32), "geom_conaffinity": G("conaffinity", -1, np.int32),
        "geom_condim": G("condim", -1, np.int32), "geom_priority": G("priority", -1, np.int32),
        "geom_plane": geom_plane,
        "geom_pos": G("pos", (ngeom, 3)), "geom_quat": G("quat", (ngeom, 4)), "geom_size": G("size", (ngeom, 3)),
        "geom_center": G("center", (ngeom, 3)),
        "geom_friction": G("friction", (ngeom, 3)), "geom_solref": G("solref", (ngeom, 2)), "geom_solimp": G("solimp", (ngeom, 5)),
        "geom_margin": G("margin", -1), "geom_gap": G("gap", -1), "geom_rbound": G("rbound", -1),
        "geom_hull_adr": G("hull_adr", -1, np.int32), "geom_hull_num": G("hull_num", -1, np.int32),
        "hull_vert": np.asarray(hull_vert, dtype=np.float32).reshape(-1, 3),
        "hull_nbr_adr": i32(hull_nbr_adr), "hull_nbr": i32(hull_nbr),
        "site_body": i32(site_body), "site_pos": f64(site_pos, (len(site_body), 3)), "site_size": f64(site_size, -1),
        "sensor_site": i32(sensor_site),
    }
    names = dict(body=body_names, joint=jnt_names, geom=geom_names, site=site_names, actuator=act_names, sensor=sensor_names)
    return CompiledModel(arrays, names)


PLANEMESH_MAXCON = 4
# Further details of MuJoCo's engine that SURVEY.md Appendix A marks as recalled, not verified -- stored in the model file so that
# a cross-check against MuJoCo (tools/mujoco_crosscheck.py) corrects a wrong guess by re-saving the model, not by editing code.
# Oracle and CUDA kernel honour every one of them (tests/test_oracle_physics.py, tests/test_gpu_physics.py at both values).
MPR_ITERATIONS = 50        # opt_int[8]   opt.mpr_iterations
PLANEMESH_ALLVERTS = 0     # opt_int[9]   extra plane-mesh contacts from 0 = hull-graph neighbours of the support vertex, 1 = all hull vertices
PLANEMESH_SEPVERT = 0      # opt_int[10]  their minimum separation measured between 0 = contact points, 1 = hull vertices
WARM_AFTER_NOSLIP = 0      # opt_int[11]  qacc_warmstart saved 0 = before the noslip pass, 1 = after
MPR_TOLERANCE = 1e-6       # opt_real[8]  opt.mpr_tolerance
PLANEMESH_SEP = 0.3        # opt_real[9]  that separation as a fraction of the geom's rbound
PYRAMID_RFAC = 2.0         # opt_real[10] R of a pyramidal contact's four edges = this * mu_reg^2 * R[first]


def load_model(path: str) -> CompiledModel:
    """Load ``.nmb`` directly, or compile ``.xml``.  If an ``.xml`` path does not exist but a
    sibling ``.nmb`` with the same stem does, the compiled file is used (the GPU box only carries
    compiled models)."""
    if path.endswith(".nmb"):
        return CompiledModel.load(path)
    if os.path.exists(path):
        return compile_mjcf(path)
    alt = os.path.splitext(path)[0] + ".nmb"
    for cand in (alt, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), alt)):
        if os.path.exists(cand):
            return CompiledModel.load(cand)
    raise MJCFError(f"model file not found: {path} (and no compiled {alt})")
