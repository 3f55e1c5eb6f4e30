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
