"""``NightmareV3Env`` — GPU-resident drop-in for the reference's vectorised hexapod environment.

Same constructor, attributes and ``reset/step/get_observations`` contract as the reference class
(``envs/nightmare_v3_env.py:26-396``), so ``train.py`` and rsl_rl's ``OnPolicyRunner`` run on it
unchanged.  Where the reference loops over ``num_envs`` CPU ``MjData`` objects and a dozen pybind
calls per env per step (``:191-226``), this class makes ONE C-ABI call per step
(``nm_step`` in ``include/nightmare_b200.h``) that launches one fused sm_100a kernel; actions,
state, observations, rewards and resets never leave HBM.

Deliberate differences (also listed in DESIGN.md):
* tensors returned by ``step`` are views of persistent device buffers (the reference allocates new
  CPU tensors every call, ``:311``); pass ``copy_outputs=True`` to get fresh copies;
* command resampling uses a counter-based Philox stream keyed by ``(seed, env id, step)`` instead
  of the unseeded global numpy RNG (``:327-330``), which cannot be reproduced;
* ``cfg.viewer.render`` is ignored with a warning (no GUI on a GPU box).
"""
from __future__ import annotations

import os
import pickle
import time
import warnings

import numpy as np
import torch

from .. import _lib, mjcf
from ..batch import Batch
from ..envcfg import REWARD_TERMS, build_envcfg, reward_table
from .nightmare_v3_config import NightmareV3Config


class NightmareV3Env:
    def __init__(self, cfg: NightmareV3Config, log_dir="/tmp/nightmare_v3/logs", num_threads=1, *,
                 device=None, seed: int = 0, env_offset: int = 0, copy_outputs: bool = False, debug: bool = False):
        self.cfg = cfg
        self.log_dir = log_dir
        self.thread_num = num_threads            # accepted for API compatibility; the GPU needs no host threads
        self.env_offset = int(env_offset)        # global id of local env 0 (multi-GPU sharding)

        self.num_envs = self.cfg.env.num_envs
        self.num_obs = self.cfg.env.num_obs
        self.num_privileged_obs = self.num_obs   # reference quirk: not None although privileged obs are None (:34)
        self.num_actions = self.cfg.env.num_actions

        if device is None:
            device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.NightmareLibError("NightmareV3Env needs a CUDA device: the environment step has no CPU path")

        # ≙ mj.MjModel.from_xml_path (:37): host compile (or load of the pre-compiled .nmb), then upload
        self.model = mjcf.load_model(self.cfg.env.model_path)
        self._dev_model = _lib.Model(self.model.to_bytes())
        self.num_dof = self.model.nv - 6
        self.num_geoms = self.model.ngeom
        self.num_oscillators = self.num_actions - self.num_dof
        self.gravity_vec = np.array([0.0, 0.0, -9.81])
        self.body_index = self.model.name2id(mjcf.OBJ_BODY, self.cfg.env.body_name)
        assert self.body_index != -1
        if self.body_index != 1:
            raise _lib.NightmareLibError("cfg.env.body_name must name the floating base (body 1)")
        if self.num_obs != 66 or self.num_dof != 18:
            raise _lib.NightmareLibError("the step kernel assembles the 66-entry observation of the 18-dof hexapod")

        if self.cfg.viewer.render:
            warnings.warn("cfg.viewer.render is ignored: the GPU environment has no interactive viewer")
        self.recorded_states = []

        self.reward_scales, self.reward_names, self._sum_keys = None, None, None
        self.dt = self.model.opt_real[0] * self.cfg.control.decimation
        table, order, keys = reward_table(self.cfg, self.dt)
        self.reward_scales = {k: float(table[REWARD_TERMS.index(k)]) for k in keys}
        self.reward_names = order
        self._sum_keys = keys
        self.command_ranges = self.cfg.commands.ranges
        self.obs_scales = self.cfg.normalization.obs_scales
        self.max_episode_length_s = self.cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.default_dof_pos = np.array(self.cfg.control.default_pos, dtype=np.float64)
        self.add_noise = self.cfg.noise.add_noise
        self.common_step_counter = 0
        self.extras = {}
        self._copy = copy_outputs
        self._host = None

        self._envcfg = build_envcfg(self.cfg, float(self.model.opt_real[0]))
        self._batch = Batch(self._dev_model, self.num_envs, self.device, seed=seed, envcfg=self._envcfg,
                            debug=debug, env_offset=env_offset)
        b = self._batch
        # public buffers (device tensors; the reference keeps numpy arrays of the same names, :56-97)
        self.obs_buf, self.rew_buf, self.reset_buf = b.obs, b.rew, b.done
        self.privileged_obs_buf = None
        self.time_out_buf = b.time_outs
        self.commands, self.actions = b.commands, b.actions
        self.dof_pos, self.dof_vel = b.dof_pos, b.dof_vel
        self.feet_air_time = b.feet_air_time
        self.episode_sums = {k: b.episode_sums[:, REWARD_TERMS.index(k)] for k in keys}
        self._extras_keys = [("rew_" + k, REWARD_TERMS.index(k)) for k in keys]
        _views = b.ep_means.unbind(0)
        self._extras_views = {name: _views[i] for name, i in self._extras_keys}
        self._rec = None
        if self.cfg.viewer.record_states:
            self._rec = _StateRecorder(self, self.log_dir)

    # ------------------------------------------------------------------ episode_length_buf: the runner REBINDS it
    @property
    def episode_length_buf(self):
        return self._batch.episode_length

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        # rsl_rl: env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=...) (train.py:54)
        self._batch.episode_length.copy_(torch.as_tensor(value).to(device=self.device, dtype=torch.int64))

    # ------------------------------------------------------------------ step
    def step(self, actions):
        """Apply actions, advance ``decimation`` physics substeps, return ``(obs, None, rew, dones, extras)``
        exactly like the reference (:145-311); every tensor is on ``self.device``."""
        self.render()
        self.common_step_counter += 1
        self._batch.step(actions, self.common_step_counter)
        self._refresh_extras()
        if self._rec is not None:
            self._rec.after_step()
        if self._copy:
            return self.obs_buf.clone(), None, self.rew_buf.clone(), self.reset_buf.clone(), self.extras
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras

    def fast_step(self, actions):
        """``step`` without building the return tuple / extras dict (pre-bound rollout loop of the PPO runner, which reads
        the env's persistent device buffers directly).  ``actions``: contiguous float32 CUDA tensor [num_envs, >=18]."""
        self.common_step_counter += 1
        self._batch.step(actions, self.common_step_counter)
        if self._rec is not None:
            self._rec.after_step()

    def step_host(self, actions, obs_out=None, rew_out=None, done_out=None):
        """``step`` for HOST-resident callers, the way the reference is used (CPU action tensor in, CPU tensors out,
        ``envs/nightmare_v3_env.py:155,311``): one C-ABI call (``nm_step_host``) enqueues the host->device copy of the
        actions, the step, and the device->host copies of obs / rew / dones on the current stream and waits for them.
        ``actions`` should be a pinned CPU float32 tensor ``[num_envs, >=18]``; outputs are written into the given
        (pinned) tensors or into buffers owned by the env, which are overwritten by the next call."""
        if self._host is None:
            pin = lambda *s, **k: torch.empty(*s, **k).pin_memory()
            self._host = (pin(self.num_envs, 66), pin(self.num_envs), pin(self.num_envs, dtype=torch.int64))
            self._host_call = _lib.lib.nm_step_host
            self._host_touts = self._batch.time_outs_latched if self.cfg.env.send_timeouts else None
        if obs_out is None:
            obs_out, rew_out, done_out = self._host
        a = actions
        if not torch.is_tensor(a) or a.device.type != "cpu" or a.dtype != torch.float32 or not a.is_contiguous():
            a = torch.as_tensor(a).detach().to("cpu", torch.float32).contiguous()
        shp = a.shape
        if len(shp) != 2 or shp[0] != self.num_envs or shp[1] < 18:
            raise ValueError(f"actions must be [num_envs, >=18], got {tuple(shp)}")
        self.common_step_counter += 1
        b = self._batch
        rc = self._host_call(b._h, a.data_ptr(), shp[1], self.common_step_counter, obs_out.data_ptr(), rew_out.data_ptr(), done_out.data_ptr(),
                             torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            _lib.check(rc)
        ex = self.extras                                      # views of the device-side latches (valid until the next step)
        ex["episode"] = self._extras_views
        if self._host_touts is not None:
            ex["time_outs"] = self._host_touts
        if self._rec is not None:
            self._rec.after_step()
        return obs_out, None, rew_out, done_out, self.extras

    def _refresh_extras(self, fresh=True):
        # extras are only refreshed on steps where at least one env reset (reference quirk Q10, :344,:363-371).
        # The latching happens on the device (nm_finalize_kernel, second launch of nm_step): no host sync.  With
        # ``fresh`` one small clone gives the dict its own storage, so dicts a runner keeps (rsl_rl appends
        # infos['episode'] every step) retain the values of THEIR step, like the reference's fresh tensors.
        if fresh:
            means = self._batch.ep_means.clone().unbind(0)
            self.extras["episode"] = {name: means[i] for name, i in self._extras_keys}
        else:
            self.extras["episode"] = self._extras_views
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self._batch.time_outs_latched

    def get_observations(self):
        return self.obs_buf.clone() if self._copy else self.obs_buf

    def get_privileged_observations(self):
        return None

    # ------------------------------------------------------------------ reset
    def reset_idx(self, env_ids):
        """≙ reference ``reset_idx`` (:335-371) for an explicit list of env ids."""
        ids = torch.as_tensor(np.asarray(env_ids) if not torch.is_tensor(env_ids) else env_ids, device=self.device).to(torch.int64).flatten()
        if ids.numel() == 0:
            return
        b = self._batch
        b.ep_means.copy_(b.episode_sums[ids].mean(dim=0) / self.max_episode_length_s)
        b.reset_idx(ids, self.common_step_counter)
        b.time_outs_latched.copy_(b.time_outs)
        self._refresh_extras()

    def reset(self):
        """Reset all robots, then take one zero-action step (:392-396)."""
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(torch.zeros((self.num_envs, self.num_actions), device=self.device))
        return obs, privileged_obs

    def render(self):
        return None

    # ------------------------------------------------------------------ domain randomisation (new, opt-in; not in the reference)
    def set_domain_randomization(self, friction=(0.5, 1.25), kv=(0.8, 1.2), base_mass=(-0.3, 0.3), resample_on_reset=True, seed=0):
        """Per-env physical parameters (BASELINE config 3): contact-friction scale ~ U(friction), actuator kv scale ~ U(kv),
        base mass + U(base_mass) kg.  Drawn now for every env and, with ``resample_on_reset``, again inside the step
        kernel whenever an env resets.  ``set_domain_randomization(None)`` switches it off.  Parity with the reference is
        defined with DR off (the reference has none)."""
        if friction is None:
            self._batch.set_domain_randomization(None)
            return None
        m0 = float(self.model.arrays["body_mass"][1])
        ranges = [friction[0], friction[1], kv[0], kv[1], 1.0 + base_mass[0] / m0, 1.0 + base_mass[1] / m0]
        g = torch.Generator(device=self.device).manual_seed(int(seed) + 7919 * self.env_offset)
        u = torch.rand(self.num_envs, 4, device=self.device, generator=g)
        lo = torch.tensor([ranges[0], ranges[2], ranges[4], 0.0], device=self.device)
        hi = torch.tensor([ranges[1], ranges[3], ranges[5], 0.0], device=self.device)
        dr = (lo + u * (hi - lo)).contiguous()
        self._batch.set_domain_randomization(dr, ranges, resample_on_reset)
        return dr

    # ------------------------------------------------------------------ raw state access (parity harness, checkpoints)
    def get_state(self):
        b = self._batch
        return b.qpos, b.qvel, b.warm

    def set_state(self, qpos=None, qvel=None, warm=None):
        b = self._batch
        for dst, src in ((b.qpos, qpos), (b.qvel, qvel), (b.warm, warm)):
            if src is not None:
                dst.copy_(torch.as_tensor(src, dtype=torch.float32).to(self.device))

    @property
    def gpu_launches(self) -> int:
        return self._batch.launches


class _StateRecorder:
    """Env-0 trajectory recorder (reference :261-272, replayed by ``open_custom_play.py:50-66``).

    The reference appends ``(time, qpos, qvel, act)`` of env 0 every step -- BEFORE ``reset_idx`` (:272 precedes :274),
    so on a reset step the row holds the terminal state -- and pickles the list whenever env 0 resets, before appending
    that step's row.  Here the step kernel itself writes env 0's pre-reset ``(done, qpos, qvel)`` into a device ring
    (``nm_batch_set_recorder``; no extra launches), which is flushed with one small D2H copy every ``flush_every``
    steps: same rows, same files (file names carry the flush time instead of the reset time)."""

    def __init__(self, env: "NightmareV3Env", log_dir: str, flush_every: int = 256):
        self.env, self.log_dir, self.k = env, log_dir, flush_every
        self.ring = torch.zeros(flush_every, _lib.NM_REC_STRIDE, device=env.device)
        env._batch.set_recorder(self.ring)
        self.fill = 0
        self.rows: list = []
        self.sim_time = 0.0

    def after_step(self):
        self.fill += 1
        if self.fill == self.k:
            self.flush()

    def flush(self):
        if self.fill == 0:
            return
        host = self.ring[: self.fill].cpu().numpy().astype(np.float64)
        self.fill = 0
        self.env._batch.set_recorder(self.ring)          # restart the ring's write cursor at row 0
        for r in host:
            self.sim_time += self.env.dt          # data.time is never reset by reset_idx (quirk Q3)
            if r[0] != 0:
                os.makedirs(self.log_dir, exist_ok=True)
                with open(f"{self.log_dir}/{int(time.time())}.pkl", "wb") as fh:
                    pickle.dump(self.rows, fh)
                self.rows = []
            self.rows.append((self.sim_time, r[1:26].copy(), r[26:50].copy(), np.zeros(0)))
        self.env.recorded_states = self.rows
