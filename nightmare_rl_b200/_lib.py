"""ctypes binding of ``libnightmare_b200.so`` (the C ABI of include/nightmare_b200.h).

There is NO CPU fallback: if the library is missing or cannot be loaded the import of this module
raises, and so does every product code path that needs it."""
from __future__ import annotations

import ctypes
import os

from .build import LIB_PATH
from .envcfg import EnvCfgStruct, NREW

NM_DBG_STRIDE = 320
NM_REC_STRIDE = 52
_vp, _ci, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64


class NmBuffers(ctypes.Structure):
    _fields_ = [(n, _vp) for n in (
        "qpos", "qvel", "warm", "actions", "dof_pos", "dof_vel", "commands", "episode_length", "episode_sums",
        "feet_air_time", "contact_bits", "obs", "rew", "done", "time_outs", "sensordata", "episode_acc", "debug",
        "ep_means", "time_outs_latched")]


class NightmareLibError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise NightmareLibError(
            f"{LIB_PATH} is missing: build it with `python -m nightmare_rl_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback for the environment step.")
    L = ctypes.CDLL(LIB_PATH)
    L.nm_last_error.restype = ctypes.c_char_p
    L.nm_model_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(_vp)]
    L.nm_model_from_buffer.argtypes = [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]
    L.nm_model_destroy.argtypes = [_vp]
    L.nm_model_size.argtypes = [_vp, ctypes.c_char_p]
    L.nm_model_timestep.argtypes = [_vp]
    L.nm_model_timestep.restype = ctypes.c_double
    L.nm_name2id.argtypes = [_vp, _ci, ctypes.c_char_p]
    L.nm_model_qpos0.argtypes = [_vp, _vp, _ci]
    L.nm_batch_create.argtypes = [_vp, _ci, _ci, ctypes.c_uint64, _vp, ctypes.POINTER(NmBuffers), ctypes.POINTER(_vp)]
    L.nm_batch_destroy.argtypes = [_vp]
    L.nm_batch_set_env_offset.argtypes = [_vp, _i64]
    L.nm_batch_set_recorder.argtypes = [_vp, _vp, _ci]
    L.nm_batch_set_domain_randomization.argtypes = [_vp, _vp, ctypes.POINTER(ctypes.c_float), _ci]
    L.nm_step.argtypes = [_vp, _vp, _ci, _i64, _vp]
    L.nm_physics_step.argtypes = [_vp, _vp, _ci, _vp]
    L.nm_reset_idx.argtypes = [_vp, _vp, _ci, _i64, _vp]
    L.nm_step_host.argtypes = [_vp, _vp, _ci, _i64, _vp, _vp, _vp, _vp]
    L.nm_batch_launches.argtypes = [_vp]
    L.nm_batch_launches.restype = _i64
    L.nm_measure_fp32_peak.argtypes = [_vp]
    L.nm_measure_fp32_peak.restype = ctypes.c_double
    L.nm_policy_create.argtypes = [_vp, _vp, _ci, ctypes.POINTER(_vp)]
    L.nm_policy_destroy.argtypes = [_vp]
    L.nm_policy_param_count.argtypes = [_vp, _ci]
    L.nm_policy_load_weights.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.nm_policy_act.argtypes = [_vp, _vp, _ci, _ci, ctypes.c_uint64, _i64, _i64, _ci, _vp, _vp, _vp, _vp, _vp]
    L.nm_policy_act_store.argtypes = [_vp, _vp, _ci, _ci, ctypes.c_uint64, _i64, _i64, _ci, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.nm_policy_launches.argtypes = [_vp]
    L.nm_policy_launches.restype = _i64
    L.nm_rollout_store.argtypes = [_vp, _vp]
    L.nm_ppo_head.argtypes = [_vp, _vp]
    L.nm_ppo_grad.argtypes = [_vp, _vp, _vp, _vp]
    L.nm_ppo_adam.argtypes = [_vp, _vp]
    L.nm_gae.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp]
    L.nm_policy_tc5_create.argtypes = [_vp, _vp, _ci, ctypes.POINTER(_vp)]
    L.nm_policy_tc5_destroy.argtypes = [_vp]
    L.nm_policy_tc5_load_weights.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.nm_policy_tc5_act.argtypes = [_vp, _vp, _ci, _ci, ctypes.c_uint64, _i64, _i64, _ci, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.nm_gen_model_from_buffer.argtypes = [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]
    L.nm_gen_model_destroy.argtypes = [_vp]
    L.nm_gen_model_size.argtypes = [_vp, ctypes.c_char_p]
    L.nm_gen_model_timestep.argtypes = [_vp]
    L.nm_gen_model_timestep.restype = ctypes.c_double
    L.nm_gen_model_qpos0.argtypes = [_vp, _vp, _ci]
    L.nm_gen_batch_create.argtypes = [_vp, _ci, _ci, _vp, _vp, _vp, _vp, ctypes.POINTER(_vp)]
    L.nm_gen_batch_destroy.argtypes = [_vp]
    L.nm_gen_physics_step.argtypes = [_vp, _vp, _ci, _vp]
    L.nm_gen_batch_launches.argtypes = [_vp]
    L.nm_gen_batch_launches.restype = _i64
    return L


lib = _load()

EXPORTS = ("nm_last_error", "nm_model_load", "nm_model_from_buffer", "nm_model_destroy", "nm_model_size", "nm_model_timestep",
           "nm_name2id", "nm_model_qpos0", "nm_model_support_map", "nm_batch_create", "nm_batch_destroy", "nm_batch_set_env_offset",
           "nm_batch_set_domain_randomization", "nm_batch_set_recorder", "nm_step",
           "nm_physics_step", "nm_reset_idx", "nm_step_host", "nm_batch_launches", "nm_measure_fp32_peak",
           "nm_policy_create", "nm_policy_destroy", "nm_policy_param_count", "nm_policy_load_weights", "nm_policy_act",
           "nm_policy_act_store", "nm_policy_launches", "nm_rollout_store", "nm_ppo_head", "nm_ppo_grad", "nm_gae", "nm_ppo_adam",
           "nm_policy_tc5_create", "nm_policy_tc5_destroy", "nm_policy_tc5_load_weights", "nm_policy_tc5_act",
           "nm_gen_model_from_buffer", "nm_gen_model_destroy", "nm_gen_model_size", "nm_gen_model_timestep", "nm_gen_model_qpos0",
           "nm_gen_batch_create", "nm_gen_batch_destroy", "nm_gen_physics_step", "nm_gen_batch_launches")


def check(rc: int) -> None:
    if rc != 0:
        raise NightmareLibError(f"nightmare_b200 error {rc}: {lib.nm_last_error().decode()}")


class Model:
    """Device-ready compiled model (≙ ``mj.MjModel``)."""

    def __init__(self, nmb_bytes: bytes):
        self._h = _vp()
        buf = ctypes.create_string_buffer(nmb_bytes, len(nmb_bytes))
        check(lib.nm_model_from_buffer(buf, len(nmb_bytes), ctypes.byref(self._h)))

    def size(self, what: str) -> int:
        return lib.nm_model_size(self._h, what.encode())

    @property
    def timestep(self) -> float:
        return lib.nm_model_timestep(self._h)

    def name2id(self, objtype: int, name: str) -> int:
        return lib.nm_name2id(self._h, objtype, name.encode())

    def qpos0(self):
        out = (ctypes.c_float * 64)()
        n = lib.nm_model_qpos0(self._h, out, 64)
        return list(out[:n])

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.nm_model_destroy(self._h)
            self._h = None


class GenModel:
    """Device-ready compiled model for the Newton / elliptic-cone path (≙ ``mj.MjModel`` of models/anymal_c)."""

    def __init__(self, nmb_bytes: bytes):
        self._h = _vp()
        buf = ctypes.create_string_buffer(nmb_bytes, len(nmb_bytes))
        check(lib.nm_gen_model_from_buffer(buf, len(nmb_bytes), ctypes.byref(self._h)))

    def size(self, what: str) -> int:
        return lib.nm_gen_model_size(self._h, what.encode())

    @property
    def timestep(self) -> float:
        return lib.nm_gen_model_timestep(self._h)

    def qpos0(self):
        out = (ctypes.c_float * 64)()
        n = lib.nm_gen_model_qpos0(self._h, out, 64)
        return list(out[:n])

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.nm_gen_model_destroy(self._h)
            self._h = None
