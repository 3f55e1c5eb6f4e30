"""ctypes front-end of the CPU oracle (oracle/nm_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by anything under nightmare_rl_b200/."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIBS = {}
# "f64": the oracle.  "f32": the same source in fp32 arithmetic (rounding floor of an fp32 implementation).
# "cnt": fp64 with floating-point operation counters (algorithmic FLOP count of one env step).
_SO = {"f64": "libnm_oracle.so", "f32": "libnm_oracle_f32.so", "cnt": "libnm_oracle_cnt.so"}


def build(force: bool = False, variant: str = "f64") -> str:
    so = os.path.join(_HERE, _SO[variant])
    src = [os.path.join(_HERE, f) for f in ("nm_oracle.c", "nm_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s", _SO[variant]])
    return so


def lib(variant: str = "f64"):
    global _LIB
    if variant not in _LIBS:
        L = ctypes.CDLL(build(variant=variant))
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
        L.nmo_model_load.restype = vp
        L.nmo_model_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ci]
        L.nmo_model_free.argtypes = [vp]
        L.nmo_model_size.argtypes = [vp, ctypes.c_char_p]
        L.nmo_batch_create.restype = vp
        L.nmo_batch_create.argtypes = [vp, ci, ctypes.c_uint64, vp]
        L.nmo_batch_free.argtypes = [vp]
        L.nmo_set_state.argtypes = [vp, vp, vp, vp]
        L.nmo_get_state.argtypes = [vp, vp, vp, vp]
        L.nmo_physics_step.argtypes = [vp, vp, ci, ci]
        L.nmo_forward.argtypes = [vp, vp, ci]
        L.nmo_get_array.argtypes = [vp, ci, ctypes.c_char_p, vp, ci]
        L.nmo_env_step.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, ci]
        L.nmo_env_reset_idx.argtypes = [vp, vp, ci]
        L.nmo_env_get.argtypes = [vp, ctypes.c_char_p, vp, ci]
        L.nmo_env_set.argtypes = [vp, ctypes.c_char_p, vp, ci]
        L.nmo_philox4x32.argtypes = [ctypes.c_uint32] * 6 + [vp]
        if variant == "cnt":
            L.nmo_flop_counts.argtypes = [vp, ci]
            L.nmo_flop_reset.argtypes = []
        _LIBS[variant] = L
        if variant == "f64":
            _LIB = L
    return _LIBS[variant]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def philox4x32(k0, k1, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().nmo_philox4x32(k0, k1, c0, c1, c2, c3, _ptr(out))
    return out


class OracleModel:
    def __init__(self, nmb_path: str, variant: str = "f64"):
        err = ctypes.create_string_buffer(256)
        self.variant = variant
        self.L = lib(variant)
        self.h = self.L.nmo_model_load(nmb_path.encode(), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensor"):
            setattr(self, k, self.L.nmo_model_size(self.h, k.encode()))

    def __del__(self):
        if getattr(self, "h", None) and getattr(self, "L", None) is not None:
            self.L.nmo_model_free(self.h)
            self.h = None


class OracleBatch:
    """N independent environments stepped on the CPU in fp64."""

    def __init__(self, model: OracleModel, num_envs: int, seed: int = 0, envcfg=None):
        self.model, self.n = model, num_envs
        self._cfg = envcfg
        self.L = model.L
        self.h = self.L.nmo_batch_create(model.h, num_envs, seed, ctypes.byref(envcfg) if envcfg is not None else None)
        if not self.h:
            raise RuntimeError("nmo_batch_create failed")

    def __del__(self):
        if getattr(self, "h", None) and getattr(self, "L", None) is not None:
            self.L.nmo_batch_free(self.h)
            self.h = None

    # ---- raw physics
    def set_state(self, qpos=None, qvel=None, warm=None):
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        qpos, qvel, warm = c(qpos), c(qvel), c(warm)
        self.L.nmo_set_state(self.h, _ptr(qpos), _ptr(qvel), _ptr(warm))

    def get_state(self):
        m = self.model
        qpos, qvel, warm = np.zeros((self.n, m.nq)), np.zeros((self.n, m.nv)), np.zeros((self.n, m.nv))
        self.L.nmo_get_state(self.h, _ptr(qpos), _ptr(qvel), _ptr(warm))
        return qpos, qvel, warm

    def physics_step(self, ctrl=None, nstep=1, nthreads=1):
        ctrl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        self.L.nmo_physics_step(self.h, _ptr(ctrl), nstep, nthreads)

    def forward(self, ctrl=None, nthreads=1):
        ctrl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        self.L.nmo_forward(self.h, _ptr(ctrl), nthreads)

    def get(self, env: int, name: str, cap: int = 65536):
        buf = np.zeros(cap)
        n = self.L.nmo_get_array(self.h, env, name.encode(), _ptr(buf), cap)
        if n < 0:
            raise KeyError(name)
        return buf[:n].copy()

    # ---- env layer
    def env_step(self, actions, nthreads=1):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        obs = np.zeros((self.n, 66), dtype=np.float32)
        rew = np.zeros(self.n, dtype=np.float32)
        done = np.zeros(self.n, dtype=np.int64)
        tout = np.zeros(self.n, dtype=np.float32)
        means = np.zeros(18)
        nres = ctypes.c_int(0)
        self.L.nmo_env_step(self.h, _ptr(a), a.shape[1], _ptr(obs), _ptr(rew), _ptr(done), _ptr(tout), _ptr(means),
                           ctypes.byref(nres), nthreads)
        return obs, rew, done, tout, means, nres.value

    def env_reset_idx(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self.L.nmo_env_reset_idx(self.h, _ptr(ids), ids.size)

    def env_get(self, name: str):
        per = {"ep_len": 1, "commands": 3, "actions": 18, "dof_pos": 18, "dof_vel": 18, "episode_sums": 18,
               "feet_air_time": 6, "last_contacts": 6, "last_contacts_filt": 6, "reset_buf": 1, "time_out": 1,
               "tibia_f": 6, "feet_f": 6, "body_f": 1, "base_lin_vel": 3, "base_ang_vel": 3, "projected_gravity": 3}.get(name)
        if name == "step_counter":
            buf = np.zeros(1)
            self.L.nmo_env_get(self.h, name.encode(), _ptr(buf), 1)
            return buf[0]
        buf = np.zeros(self.n * per)
        if self.L.nmo_env_get(self.h, name.encode(), _ptr(buf), buf.size) < 0:
            raise KeyError(name)
        return buf.reshape(self.n, per) if per > 1 else buf

    def env_set(self, name: str, values):
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        rc = self.L.nmo_env_set(self.h, name.encode(), _ptr(v), v.size)
        if rc != 0:
            raise ValueError(f"env_set({name}) failed: {rc}")
