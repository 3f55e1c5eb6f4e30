"""Device-resident batch of environments: torch tensors for storage, C-ABI calls for compute."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .envcfg import NREW, EnvCfgStruct


class Batch:
    """≙ ``[mj.MjData(model) for _ in range(num_envs)]`` (reference envs/nightmare_v3_env.py:38), in HBM.

    All state lives in torch CUDA tensors (row-major ``[num_envs, K]`` fp32); the CUDA library only
    receives their raw pointers.  ``envcfg=None`` creates a physics-only batch (parity harness)."""

    def __init__(self, model: _lib.Model, num_envs: int, device: torch.device, seed: int = 0,
                 envcfg: EnvCfgStruct | None = None, debug: bool = False, env_offset: int = 0):
        if device.type != "cuda":
            raise _lib.NightmareLibError("the environment step only runs on CUDA devices (no CPU fallback)")
        self.model, self.n, self.device = model, num_envs, device
        f32 = dict(dtype=torch.float32, device=device)
        z = lambda *s: torch.zeros(*s, **f32)
        self.qpos = torch.tensor(model.qpos0(), **f32).repeat(num_envs, 1).contiguous()
        self.qvel, self.warm = z(num_envs, 24), z(num_envs, 24)
        self.actions, self.dof_pos, self.dof_vel = z(num_envs, 18), z(num_envs, 18), z(num_envs, 18)
        self.commands = z(num_envs, 3)
        self.episode_length = torch.zeros(num_envs, dtype=torch.int64, device=device)
        self.episode_sums = z(num_envs, NREW)
        self.feet_air_time = z(num_envs, 6)
        self.contact_bits = torch.zeros(num_envs, dtype=torch.int32, device=device)
        self.obs, self.rew = z(num_envs, 66), z(num_envs)
        self.done = torch.ones(num_envs, dtype=torch.int64, device=device)
        self.time_outs = z(num_envs)
        self.sensordata = z(num_envs, 13)
        self.episode_acc = z(NREW + 1)
        self.ep_means, self.time_outs_latched = z(NREW), z(num_envs)
        self.debug = z(num_envs, _lib.NM_DBG_STRIDE) if debug else None
        b = _lib.NmBuffers()
        for name, _ in _lib.NmBuffers._fields_:
            t = getattr(self, name)
            setattr(b, name, None if t is None else t.data_ptr())
        self._bufs = b
        self._cfg = envcfg
        self._h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib.nm_batch_create(model._h, num_envs, device.index or 0, seed,
                                                ctypes.byref(envcfg) if envcfg is not None else None,
                                                ctypes.byref(b), ctypes.byref(self._h)))
        if env_offset:
            _lib.check(_lib.lib.nm_batch_set_env_offset(self._h, env_offset))

    def set_domain_randomization(self, dr, ranges=None, resample_on_reset=False):
        """``dr``: float32 CUDA tensor [num_envs, 4] (friction, kv, base-mass scale, unused) or None to switch DR off."""
        if dr is None:
            _lib.check(_lib.lib.nm_batch_set_domain_randomization(self._h, None, None, 0))
            self.dr = None
            return
        if dr.shape != (self.n, 4) or dr.dtype != torch.float32 or dr.device != self.device or not dr.is_contiguous():
            raise ValueError("dr must be a contiguous float32 [num_envs, 4] tensor on the batch's device")
        r = (ctypes.c_float * 6)(*([float(x) for x in ranges] if ranges is not None else [1.0] * 6))
        _lib.check(_lib.lib.nm_batch_set_domain_randomization(self._h, dr.data_ptr(), r, 1 if (resample_on_reset and ranges is not None) else 0))
        self.dr = dr

    def set_recorder(self, ring):
        """``ring``: float32 CUDA tensor [capacity, NM_REC_STRIDE]; every following step writes env 0's pre-reset
        (done, qpos, qvel) into row (steps since this call) % capacity.  None switches the recorder off."""
        if ring is None:
            _lib.check(_lib.lib.nm_batch_set_recorder(self._h, None, 0))
        else:
            if ring.dim() != 2 or ring.shape[1] != _lib.NM_REC_STRIDE or ring.dtype != torch.float32 or ring.device != self.device or not ring.is_contiguous():
                raise ValueError("recorder ring must be a contiguous float32 [capacity, NM_REC_STRIDE] tensor on the batch's device")
            _lib.check(_lib.lib.nm_batch_set_recorder(self._h, ring.data_ptr(), ring.shape[0]))
        self._rec_ring = ring

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, actions: torch.Tensor, step_counter: int) -> None:
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        if actions.dim() != 2 or actions.shape[0] != self.n or actions.shape[1] < 18:
            raise ValueError(f"actions must be [num_envs, >=18], got {tuple(actions.shape)}")
        _lib.check(_lib.lib.nm_step(self._h, actions.data_ptr(), actions.shape[1], step_counter, self._stream()))
        self._keep = actions

    def physics_step(self, ctrl: torch.Tensor, nstep: int = 1) -> None:
        ctrl = ctrl.to(device=self.device, dtype=torch.float32).contiguous()
        if ctrl.shape != (self.n, 18):
            raise ValueError("ctrl must be [num_envs, 18]")
        _lib.check(_lib.lib.nm_physics_step(self._h, ctrl.data_ptr(), nstep, self._stream()))
        self._keep = ctrl

    def reset_idx(self, env_ids: torch.Tensor, step_counter: int) -> None:
        ids = env_ids.to(device=self.device, dtype=torch.int64).contiguous()
        if ids.numel():
            _lib.check(_lib.lib.nm_reset_idx(self._h, ids.data_ptr(), ids.numel(), step_counter, self._stream()))
        self._keep = ids

    def step_host(self, h_actions: torch.Tensor, step_counter: int, h_obs, h_rew, h_done) -> None:
        """End-to-end step on HOST (pinned) buffers: H2D, kernel, D2H, stream sync — all inside the library."""
        _lib.check(_lib.lib.nm_step_host(self._h, h_actions.data_ptr(), h_actions.shape[1], step_counter,
                                         h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr(), self._stream()))

    @property
    def launches(self) -> int:
        return int(_lib.lib.nm_batch_launches(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib.nm_batch_destroy(self._h)
            self._h = None


class GenBatch:
    """≙ ``[mj.MjData(model) for _ in range(num_envs)]`` for a model on the Newton / elliptic-cone path (models/anymal_c of the
    reference, BASELINE configs[3]); ``physics_step`` ≙ ``mj.mj_step(model, data[i], nstep)`` for every environment."""

    def __init__(self, model: _lib.GenModel, num_envs: int, device: torch.device):
        if device.type != "cuda":
            raise _lib.NightmareLibError("the physics step only runs on CUDA devices (no CPU fallback)")
        self.model, self.n, self.device = model, num_envs, device
        self.nq, self.nv, self.nu = model.size("nq"), model.size("nv"), model.size("nu")
        f32 = dict(dtype=torch.float32, device=device)
        self.qpos = torch.tensor(model.qpos0(), **f32).repeat(num_envs, 1).contiguous()
        self.qvel, self.warm = torch.zeros(num_envs, self.nv, **f32), torch.zeros(num_envs, self.nv, **f32)
        self.info = torch.zeros(num_envs, 4, dtype=torch.int32, device=device)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib.nm_gen_batch_create(model._h, num_envs, device.index or 0, self.qpos.data_ptr(), self.qvel.data_ptr(),
                                                    self.warm.data_ptr(), self.info.data_ptr(), ctypes.byref(self._h)))

    def physics_step(self, ctrl: torch.Tensor, nstep: int = 1) -> None:
        ctrl = ctrl.to(device=self.device, dtype=torch.float32).contiguous()
        if ctrl.shape != (self.n, self.nu):
            raise ValueError(f"ctrl must be [num_envs, {self.nu}]")
        _lib.check(_lib.lib.nm_gen_physics_step(self._h, ctrl.data_ptr(), nstep,
                                                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        self._keep = ctrl

    @property
    def launches(self) -> int:
        return int(_lib.lib.nm_gen_batch_launches(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib.nm_gen_batch_destroy(self._h)
            self._h = None
