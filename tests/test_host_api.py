"""Host-side logic and the C-ABI library surface (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, NMB


def test_library_exports_every_declared_symbol():
    from nightmare_rl_b200 import _lib
    header = open(os.path.join(ROOT, "include", "nightmare_b200.h")).read()
    declared = set(re.findall(r"\b(nm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(_lib.lib, sym), sym


def test_struct_layouts_agree():
    """nm_envcfg (include/nightmare_b200.h), nmo_envcfg (oracle) and EnvCfgStruct (ctypes) are one layout."""
    from nightmare_rl_b200.envcfg import EnvCfgStruct
    n_i32, n_f64 = 8, 4 + 18 + 4 + 2 + 2 + 3 + 3 + 1 + 18 + 66
    assert ctypes.sizeof(EnvCfgStruct) == 4 * n_i32 + 8 * n_f64
    from nightmare_rl_b200 import _lib
    assert ctypes.sizeof(_lib.NmBuffers) == 20 * ctypes.sizeof(ctypes.c_void_p)      # pointer fields of nm_buffers


def test_model_through_abi(compiled_model):
    from nightmare_rl_b200 import _lib
    m = _lib.Model(compiled_model.to_bytes())
    assert (m.size("nq"), m.size("nv"), m.size("nu"), m.size("nleg"), m.size("nsensor")) == (25, 24, 18, 6, 13)
    assert m.timestep == 0.008
    assert m.name2id(1, "base_link") == 1 and m.name2id(1, "missing") == -1      # mjOBJ_BODY = 1
    assert np.allclose(m.qpos0()[:7], [0, 0, 0.15, 1, 0, 0, 0])
    h = ctypes.c_void_p()
    assert _lib.lib.nm_model_load(NMB.encode(), ctypes.byref(h)) == 0
    assert _lib.lib.nm_model_size(h, b"nbody") == 20
    _lib.lib.nm_model_destroy(h)


def test_abi_error_codes(compiled_model, tmp_path):
    from nightmare_rl_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.lib.nm_model_load(str(tmp_path / "nope.nmb").encode(), ctypes.byref(h)) == -1      # NM_ERR_IO
    assert b"cannot open" in _lib.lib.nm_last_error()
    with pytest.raises(_lib.NightmareLibError):
        _lib.Model(b"garbage-not-nmb")
    # a model outside the kernel's topology is refused loudly, not approximated
    from nightmare_rl_b200 import mjcf
    cm = mjcf.CompiledModel(dict((k, v.copy()) for k, v in compiled_model.arrays.items()), compiled_model.names)
    cm.arrays["opt_int"][1] = mjcf.SOL_NEWTON
    with pytest.raises(_lib.NightmareLibError, match="PGS"):
        _lib.Model(cm.to_bytes())
    cm.arrays["opt_int"][1] = mjcf.SOL_PGS
    cm.arrays["jnt_pos"][3, 0] = 0.01
    with pytest.raises(_lib.NightmareLibError, match="anchors"):
        _lib.Model(cm.to_bytes())


def test_env_requires_cuda():
    """No CPU fallback: constructing the env without a GPU fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nightmare_rl_b200 import _lib
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    from nightmare_rl_b200.envs.nightmare_v3_env import NightmareV3Env
    cfg = NightmareV3Config()
    cfg.env.num_envs = 4
    cfg.env.model_path = NMB
    with pytest.raises(_lib.NightmareLibError):
        NightmareV3Env(cfg)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nightmare_rl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".hpp", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "nm_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_config_mirrors_reference_values():
    from nightmare_rl_b200.envs.helpers import class_to_dict
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
    d = class_to_dict(NightmareV3Config())
    assert d["env"]["num_obs"] == 66 and d["env"]["num_actions"] == 18 and d["control"]["decimation"] == 2
    assert d["control"]["p_gain"] == 20 and d["control"]["action_scale"] == 0.2
    assert np.allclose(d["control"]["default_pos"], [0, np.pi / 5, 0] * 6)
    assert list(d["rewards"]["scales"].keys()) == sorted(d["rewards"]["scales"].keys())       # dir() order = alphabetical
    p = class_to_dict(NightmareV3ConfigPPO())
    assert p["policy"]["actor_hidden_dims"] == [54, 42, 30] and p["runner"]["num_steps_per_env"] == 80
    if os.path.isdir("/root/reference/envs"):
        import importlib.util
        import sys
        sys.path.insert(0, "/root/reference")
        try:
            for k in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
                del sys.modules[k]
            spec = importlib.util.spec_from_file_location("ref_cfg", "/root/reference/envs/nightmare_v3_config.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            assert class_to_dict(mod.NightmareV3Config()) == d
            assert class_to_dict(mod.NightmareV3ConfigPPO()) == p
        finally:
            sys.path.remove("/root/reference")
            for k in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
                del sys.modules[k]


def test_base_config_instances_are_independent():
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config
    a, b = NightmareV3Config(), NightmareV3Config()
    a.env.num_envs = 7
    assert b.env.num_envs == 8192 and NightmareV3Config.env.num_envs == 8192


def test_get_load_path(tmp_path):
    from nightmare_rl_b200.envs.helpers import get_load_path
    for run in ("2024-03-09 10:00:00", "2024-03-10 09:00:00"):
        os.makedirs(tmp_path / run)
    for it in (50, 100, 2000):
        (tmp_path / "2024-03-10 09:00:00" / f"model_{it}.pt").write_bytes(b"")
    assert get_load_path(str(tmp_path)).endswith("2024-03-10 09:00:00/model_2000.pt")
    assert get_load_path(str(tmp_path), checkpoint=100).endswith("model_100.pt")
    with pytest.raises(ValueError):
        get_load_path(str(tmp_path / "void"))


def test_second_model_through_compiler_and_loader():
    """BASELINE configs[3] names anymal_c as the model that exercises the generic compiler / loader.  Its MJCF (includes,
    nested default classes, explicit inertials, primitive collision geoms, position actuators; reference
    models/anymal_c/anymal_c.xml) compiles to models/anymal_c/anymal_c.nmb in the same container format as the hexapod's.
    Its step needs the Newton solver / elliptic cones / condim 6 / joint limits / frictionloss: the hexapod loader
    nm_model_from_buffer must refuse it and name the loader that takes it (nm_gen_model_from_buffer, csrc/nm_generic.cu;
    GPU parity in tests/test_gpu_anymal.py), which validates the model on the host without needing a GPU."""
    from conftest import ROOT
    from nightmare_rl_b200 import _lib, mjcf
    path = os.path.join(ROOT, "models", "anymal_c", "anymal_c.nmb")
    cm = mjcf.CompiledModel.load(path)
    assert (cm.nq, cm.nv, cm.nu, cm.nbody, cm.ngeom) == (19, 18, 12, 14, 45)
    assert cm.names["body"][:5] == ["world", "base", "LF_HIP", "LF_THIGH", "LF_SHANK"]
    assert abs(cm.arrays["body_mass"].sum() - 44.96518) < 1e-6
    assert np.bincount(cm.arrays["geom_type"], minlength=8).tolist() == [1, 0, 4, 0, 0, 29, 11, 0]      # plane, 4 foot spheres, cylinders, boxes
    assert abs(float(cm.arrays["opt_real"][0]) - 0.002) < 1e-12
    # the C ABI validates a model against what the step kernels implement when it builds the device image: refused, with the reason
    with pytest.raises(_lib.NightmareLibError) as ei:
        _lib.Model(cm.to_bytes())
    assert 'only solver="PGS" is implemented' in str(ei.value) and "nm_gen_model_from_buffer" in str(ei.value)
    gm = _lib.GenModel(cm.to_bytes())                              # host-side validation + packing only: no CUDA call
    assert (gm.size("nq"), gm.size("nv"), gm.size("nu"), gm.size("nbody"), gm.size("ngeom")) == (19, 18, 12, 14, 44)
    assert abs(gm.timestep - 0.002) < 1e-9 and np.allclose(gm.qpos0(), cm.qpos0)
    with pytest.raises(_lib.NightmareLibError, match="Newton"):
        _lib.GenModel(open(os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb"), "rb").read())
    if os.path.isdir("/root/reference/models/anymal_c"):         # the committed file is what the compiler produces from the reference MJCF
        fresh = mjcf.compile_mjcf("/root/reference/models/anymal_c/scene.xml")
        assert set(fresh.arrays) == set(cm.arrays)
        for k, v in fresh.arrays.items():
            assert np.array_equal(v, cm.arrays[k]), k


def test_support_map_is_a_tight_upper_bound_of_the_hull_support():
    """The tibia-tibia candidate filter of the step kernel (nm_kernels.cu, `smap_support`) may only reject a pair that MPR's own
    first step (exact separating-axis test along the line between the hull centres, oracle/nm_oracle.c `mpr_penetration`) would
    reject too: the interpolated support-map value must never be below max_v d.v, for any direction, on every leg hull of the
    shipped model -- and it should be tight enough to be useful (mean slack well under a millimetre)."""
    from nightmare_rl_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.lib.nm_model_load(NMB.encode(), ctypes.byref(h)) == 0
    fn = _lib.lib.nm_model_support_map
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(5)
    D = rng.normal(size=(40000, 3)).astype(np.float32)
    D[:6] = np.vstack([np.eye(3), -np.eye(3)])                     # face centres
    D[6:14] = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])      # cube corners
    D[14:20] = [[1, 1, 0], [1, 0, 1], [0, 1, 1], [-1, 1, 0], [1, 0, -1], [0, -1, 1]]               # face edges (ties of the face choice)
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    assert fn(h, 6, None, 0, None, 0, None) == 0                   # lane 6 is the base hull: no map
    for leg in range(6):
        nv = ctypes.c_int(0)
        n = fn(h, leg, None, 0, None, 0, ctypes.byref(nv))
        assert n == 16 and 100 < nv.value < 512
        T = np.zeros(6 * (n + 1) * (n + 1), dtype=np.float32)
        V = np.zeros((nv.value, 3), dtype=np.float32)
        assert fn(h, leg, T.ctypes.data, T.size, V.ctypes.data, nv.value, ctypes.byref(nv)) == n
        T = T.reshape(6, n + 1, n + 1)
        exact = (D.astype(np.float64) @ V.astype(np.float64).T).max(axis=1)
        # the kernel's lookup, in fp32 like the kernel
        a = np.abs(D)
        f0 = (a[:, 0] >= a[:, 1]) & (a[:, 0] >= a[:, 2])
        f1 = ~f0 & (a[:, 1] >= a[:, 2])
        ax = np.where(f0, 0, np.where(f1, 1, 2))
        idx = np.arange(len(D))
        m = D[idx, ax]
        u, v = D[idx, (ax + 1) % 3], D[idx, (ax + 2) % 3]
        am = np.abs(m)
        s = (np.float32(1.0) / am).astype(np.float32)
        face = 2 * ax + (m < 0)
        gu = np.clip((u * s + np.float32(1)) * np.float32(0.5 * n), 0, n).astype(np.float32)
        gv = np.clip((v * s + np.float32(1)) * np.float32(0.5 * n), 0, n).astype(np.float32)
        iu, iv = np.minimum(gu.astype(np.int32), n - 1), np.minimum(gv.astype(np.int32), n - 1)
        fu, fv = gu - iu, gv - iv
        t00, t10, t01, t11 = T[face, iv, iu], T[face, iv, iu + 1], T[face, iv + 1, iu], T[face, iv + 1, iu + 1]
        lo = (fu * (t10 - t00) + t00).astype(np.float32)
        hi = (fu * (t11 - t01) + t01).astype(np.float32)
        bound = ((fv * (hi - lo) + lo) * am).astype(np.float32)
        slack = bound.astype(np.float64) - exact
        assert slack.min() > 0.0, (leg, slack.min())              # never below the hull's support: conservative
        assert slack.mean() < 5e-4 and slack.max() < 8e-3, (leg, slack.mean(), slack.max())
    _lib.lib.nm_model_destroy(h)
