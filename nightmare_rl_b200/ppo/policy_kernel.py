"""ctypes wrapper of the fused policy forward (``nm_policy_*`` in include/nightmare_b200.h, csrc/nm_policy.cu)."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib


class MlpShape(ctypes.Structure):
    _fields_ = [("num_layers", ctypes.c_int32), ("dims", ctypes.c_int32 * 7)]


def _shape(dims):
    s = MlpShape()
    s.num_layers = len(dims) - 1
    for i, d in enumerate(dims):
        s.dims[i] = int(d)
    return s


class FusedPolicy:
    """mean/value/actions/log-prob of an ``ActorCritic`` for a whole batch of observations in one kernel launch."""

    def __init__(self, actor_critic, device: torch.device, seed: int = 0, env_offset: int = 0, engine: str | None = None):
        if device.type != "cuda":
            raise _lib.NightmareLibError("the fused policy kernel only runs on CUDA devices")
        if getattr(actor_critic, "activation_name", "elu") != "elu":
            raise _lib.NightmareLibError("the fused policy kernel implements ELU hidden activations only")
        adims, cdims = actor_critic.layer_dims()
        self.device, self.seed, self.env_offset = device, int(seed), int(env_offset)
        self.num_actions = adims[-1]
        self._h = ctypes.c_void_p()
        a, c = _shape(adims), _shape(cdims)
        # engine: "tc5" = tcgen05 + TMEM (csrc/nm_policy_tc5.cu), "mma" = warp-level mma.sync (csrc/nm_policy.cu, any width
        # up to 128).  Default: tc5 when the network fits its limits.
        import os
        engine = engine or os.environ.get("NM_POLICY_ENGINE", "tc5")
        fits = adims[0] <= 72 and max(adims[1:-1] + cdims[1:-1] + [1]) <= 64 and adims[-1] <= 32 and cdims[-1] <= 16
        self.engine = "tc5" if (engine == "tc5" and fits) else "mma"
        L = _lib.lib
        self._fn = ((L.nm_policy_tc5_create, L.nm_policy_tc5_destroy, L.nm_policy_tc5_load_weights, L.nm_policy_tc5_act)
                    if self.engine == "tc5" else (L.nm_policy_create, L.nm_policy_destroy, L.nm_policy_load_weights, L.nm_policy_act_store))
        self._launches = 0
        with torch.cuda.device(device):
            _lib.check(self._fn[0](ctypes.byref(a), ctypes.byref(c), device.index or 0, ctypes.byref(self._h)))
        self._out = {}
        self.load(actor_critic)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load(self, actor_critic):
        """Re-pack the module's current parameters (call after every optimiser step that precedes a rollout)."""
        fa, fc = actor_critic.flat_params()
        std = actor_critic.std.detach().float().contiguous()
        _lib.check(self._fn[2](self._h, fa.data_ptr(), fc.data_ptr(), std.data_ptr(), self._stream()))
        self._std = std.clone()
        self._keep = (fa, fc, std)

    def act(self, obs: torch.Tensor, step: int, deterministic: bool = False, out=None, obs_copy=None, sigma_out=None):
        """→ (actions [n,A], mean [n,A], value [n], log_prob [n]); fresh tensors every call, or written into the
        contiguous float32 tensors given as ``out=(actions, mean, value, log_prob)`` (rows of the rollout buffer)."""
        if obs.device != self.device or obs.dtype != torch.float32:
            obs = obs.to(device=self.device, dtype=torch.float32)
        if obs.dim() != 2 or obs.stride(1) != 1:
            obs = obs.reshape(obs.shape[0], -1).contiguous()
        n = obs.shape[0]
        A = self.num_actions
        if out is not None:
            actions, mean, value, logp = out
        else:
            actions = torch.empty(n, A, device=self.device)
            mean = torch.empty(n, A, device=self.device)
            value = torch.empty(n, device=self.device)
            logp = torch.empty(n, device=self.device)
        self._launches += 1
        _lib.check(self._fn[3](self._h, obs.data_ptr(), obs.stride(0), n, ctypes.c_uint64(self.seed), ctypes.c_int64(step),
                                                ctypes.c_int64(self.env_offset), 1 if deterministic else 0, actions.data_ptr(), mean.data_ptr(),
                                                value.data_ptr(), logp.data_ptr(), None if obs_copy is None else obs_copy.data_ptr(),
                                                None if sigma_out is None else sigma_out.data_ptr(), self._stream()))
        self._keep_obs = obs
        return actions, mean, value, logp

    @property
    def std(self):
        return self._std

    @property
    def launches(self) -> int:
        return self._launches

    def __del__(self):
        if getattr(self, "_h", None) and getattr(_lib, "lib", None) is not None:
            self._fn[1](self._h)
            self._h = None


class RolloutSlot(ctypes.Structure):
    """``nm_rollout_slot`` of include/nightmare_b200.h."""
    _fields_ = ([("n", ctypes.c_int32), ("obs_dim", ctypes.c_int32), ("act_dim", ctypes.c_int32), ("ring_cap", ctypes.c_int32),
                 ("gamma", ctypes.c_float), ("pad0", ctypes.c_float)]
                + [(k, ctypes.c_void_p) for k in ("obs", "actions", "mean", "std", "value", "logp", "rew", "done", "time_outs",
                                                  "s_obs", "s_actions", "s_mu", "s_sigma", "s_values", "s_logp", "s_rewards", "s_dones",
                                                  "cur_rew", "cur_len", "ring_rew", "ring_len", "ring_count", "ep_means", "ep_acc")]
                + [("n_ep", ctypes.c_int32), ("pad1", ctypes.c_int32)])


def rollout_store(slot: RolloutSlot, device: torch.device):
    _lib.check(_lib.lib.nm_rollout_store(ctypes.byref(slot), ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))


class PPOHeadArgs(ctypes.Structure):
    """``nm_ppo_head_args`` of include/nightmare_b200.h."""
    _fields_ = ([("n", ctypes.c_int32), ("act_dim", ctypes.c_int32), ("use_clipped_value_loss", ctypes.c_int32), ("pad0", ctypes.c_int32),
                 ("clip", ctypes.c_float), ("value_coef", ctypes.c_float), ("pad1", ctypes.c_float), ("pad2", ctypes.c_float)]
                + [(k, ctypes.c_void_p) for k in ("mu", "value", "std", "actions", "old_logp", "old_mu", "old_sigma", "adv", "ret", "tgt_val",
                                                  "out", "g_mu", "g_value", "g_std")])


class FusedPPOHead(torch.autograd.Function):
    """loss = mean(clipped surrogate) + value_coef * mean(clipped value loss) - entropy_coef * mean(entropy), its value-loss /
    surrogate / KL means, and its gradients w.r.t. (mu, value, std) from ONE kernel launch (csrc/nm_policy.cu,
    nm_ppo_head) instead of ~120 element-wise autograd kernels.  Returns (loss, value_loss, surrogate_loss, kl_mean)."""

    @staticmethod
    def forward(ctx, mu, value, std, actions, old_logp, old_mu, old_sigma, adv, ret, tgt_val, clip, value_coef, entropy_coef, use_clipped):
        n, A = mu.shape
        c = lambda t: t.contiguous()
        mu_, value_, std_ = c(mu), c(value.reshape(-1)), c(std)
        bufs = [c(actions), c(old_logp.reshape(-1)), c(old_mu), c(old_sigma), c(adv.reshape(-1)), c(ret.reshape(-1)), c(tgt_val.reshape(-1))]
        out = torch.empty(3, device=mu.device)
        g_mu = torch.empty_like(mu_)
        g_value = torch.empty_like(value_)
        g_std = torch.empty_like(std_)
        h = PPOHeadArgs()
        h.n, h.act_dim, h.use_clipped_value_loss = n, A, 1 if use_clipped else 0
        h.clip, h.value_coef = float(clip), float(value_coef)
        h.mu, h.value, h.std = mu_.data_ptr(), value_.data_ptr(), std_.data_ptr()
        (h.actions, h.old_logp, h.old_mu, h.old_sigma, h.adv, h.ret, h.tgt_val) = [t.data_ptr() for t in bufs]
        h.out, h.g_mu, h.g_value, h.g_std = out.data_ptr(), g_mu.data_ptr(), g_value.data_ptr(), g_std.data_ptr()
        _lib.check(_lib.lib.nm_ppo_head(ctypes.byref(h), ctypes.c_void_p(torch.cuda.current_stream(mu.device).cuda_stream)))
        sums = out / n
        entropy = (0.5 + 0.5 * 1.8378770664093453 + torch.log(std_)).sum()           # identical for every sample
        loss = sums[0] + value_coef * sums[1] - entropy_coef * entropy
        g_std = g_std - entropy_coef / std_
        ctx.save_for_backward(g_mu, g_value, g_std)
        ctx.vshape = value.shape
        ctx._keep = (mu_, value_, std_, bufs)
        return loss, sums[1], sums[0], sums[2]

    @staticmethod
    def backward(ctx, g_loss, g_v, g_s, g_k):
        g_mu, g_value, g_std = ctx.saved_tensors
        return (g_loss * g_mu, (g_loss * g_value).view(ctx.vshape), g_loss * g_std) + (None,) * 11
