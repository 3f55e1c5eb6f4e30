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


class PPOGradArgs(ctypes.Structure):
    """``nm_ppo_grad_args`` of include/nightmare_b200.h."""
    _fields_ = ([("n", ctypes.c_int32), ("obs_dim", ctypes.c_int32), ("act_dim", ctypes.c_int32), ("use_clipped_value_loss", ctypes.c_int32),
                 ("clip", ctypes.c_float), ("value_coef", ctypes.c_float), ("entropy_coef", ctypes.c_float), ("pad0", ctypes.c_float)]
                + [(k, ctypes.c_void_p) for k in ("idx", "obs", "critic_obs", "actions", "old_logp", "old_mu", "old_sigma", "adv", "ret", "tgt_val",
                                                  "actor_params", "critic_params", "std", "g_actor", "g_critic", "g_std", "out")])


class PPOAdamArgs(ctypes.Structure):
    """``nm_ppo_adam_args`` of include/nightmare_b200.h."""
    _fields_ = ([("n_params", ctypes.c_int32), ("n_samples", ctypes.c_int32), ("adaptive", ctypes.c_int32), ("pad0", ctypes.c_int32)]
                + [(k, ctypes.c_float) for k in ("desired_kl", "max_grad_norm", "beta1", "beta2", "eps", "pad1", "pad2", "pad3")]
                + [(k, ctypes.c_void_p) for k in ("params", "grads", "exp_avg", "exp_avg_sq", "step", "lr", "sums", "loss_acc")])


class FusedPPOGrad:
    """Gradient of the PPO mini-batch loss for every parameter of an ``ActorCritic`` in one kernel launch (nm_ppo_grad).

    The module's parameters are re-seated as views of ONE flat buffer (std | actor | critic, PyTorch order) and their ``.grad``
    as views of a second one, so the kernel, the gradient all-reduce, the norm clip and the optimiser all work on the same
    memory without copies."""

    def __init__(self, actor_critic, device):
        if device.type != "cuda":
            raise _lib.NightmareLibError("nm_ppo_grad only runs on CUDA devices")
        if getattr(actor_critic, "activation_name", "elu") != "elu":
            raise _lib.NightmareLibError("nm_ppo_grad implements ELU hidden activations only")
        import torch.nn as nn
        ac = actor_critic
        adims, cdims = ac.layer_dims()
        self.shape_a, self.shape_c = _shape(adims), _shape(cdims)
        self.obs_dim, self.act_dim = adims[0], adims[-1]
        groups = [[ac.std],
                  [p for m in ac.actor if isinstance(m, nn.Linear) for p in (m.weight, m.bias)],
                  [p for m in ac.critic if isinstance(m, nn.Linear) for p in (m.weight, m.bias)]]
        listed = [p for grp in groups for p in grp]
        if {id(p) for p in listed} != {id(p) for p in ac.parameters()} or any(p.dtype != torch.float32 for p in listed):
            raise _lib.NightmareLibError("nm_ppo_grad: unexpected parameter set (std + Linear layers in float32 expected)")
        total = sum(p.numel() for p in listed)
        self.flat = torch.empty(total, device=device)
        # gradient vector and the kernel's 4 loss sums in ONE buffer: a multi-GPU job averages both with a single all-reduce
        self.ext = torch.zeros(total + 4, device=device)
        self.flat_grad = self.ext[:total]
        self.offsets = []
        self.param_slices = []                        # (parameter, offset, numel) in flat order
        # Adam moments and step count, flat like the parameters; the torch optimiser's per-parameter state entries are views
        # of them (seat_optimizer_state), so optimizer.state_dict() / load_state_dict() keep rsl_rl's checkpoint format
        self.exp_avg = torch.zeros(total, device=device)
        self.exp_avg_sq = torch.zeros(total, device=device)
        self.step = torch.zeros((), device=device)
        off = 0
        with torch.no_grad():
            for grp in groups:
                self.offsets.append(off)
                for p in grp:
                    n = p.numel()
                    self.param_slices.append((p, off, n))
                    self.flat[off:off + n].copy_(p.detach().reshape(-1))
                    p.data = self.flat[off:off + n].view_as(p)
                    p.grad = self.flat_grad[off:off + n].view_as(p)
                    off += n
        self.out = self.ext[total:]
        self.device = device
        self._keep = None

    def __call__(self, n, idx, obs, critic_obs, actions, old_logp, old_mu, old_sigma, adv, ret, tgt_val, clip, value_coef, entropy_coef,
                 use_clipped_value_loss):
        """All tensors flat rollout buffers ([rows, ...], float32, contiguous); idx int64 [n] or None.  Leaves the gradient in
        ``flat_grad`` (and therefore in every ``p.grad``) and returns ``out`` = [sum surrogate, sum value loss, sum KL, 0]."""
        for t_ in (obs, critic_obs, actions, old_logp, old_mu, old_sigma, adv, ret, tgt_val):
            if t_.dtype != torch.float32 or not t_.is_contiguous() or t_.device != self.device:
                raise _lib.NightmareLibError("nm_ppo_grad: rollout buffers must be contiguous float32 tensors on the policy's device")
        if idx is not None and (idx.dtype != torch.int64 or not idx.is_contiguous()):
            raise _lib.NightmareLibError("nm_ppo_grad: idx must be a contiguous int64 tensor")
        a = PPOGradArgs()
        a.n, a.obs_dim, a.act_dim, a.use_clipped_value_loss = int(n), self.obs_dim, self.act_dim, 1 if use_clipped_value_loss else 0
        a.clip, a.value_coef, a.entropy_coef = float(clip), float(value_coef), float(entropy_coef)
        a.idx = idx.data_ptr() if idx is not None else None
        a.obs, a.critic_obs, a.actions = obs.data_ptr(), critic_obs.data_ptr(), actions.data_ptr()
        a.old_logp, a.old_mu, a.old_sigma = old_logp.data_ptr(), old_mu.data_ptr(), old_sigma.data_ptr()
        a.adv, a.ret, a.tgt_val = adv.data_ptr(), ret.data_ptr(), tgt_val.data_ptr()
        o_std, o_a, o_c = self.offsets
        base, gbase = self.flat.data_ptr(), self.flat_grad.data_ptr()
        a.std, a.actor_params, a.critic_params = base + 4 * o_std, base + 4 * o_a, base + 4 * o_c
        a.g_std, a.g_actor, a.g_critic = gbase + 4 * o_std, gbase + 4 * o_a, gbase + 4 * o_c
        a.out = self.out.data_ptr()
        _lib.check(_lib.lib.nm_ppo_grad(ctypes.byref(self.shape_a), ctypes.byref(self.shape_c), ctypes.byref(a),
                                        ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return self.out

    def seat_optimizer_state(self, optimizer):
        """Make ``optimizer.state[p]`` = views of the flat moment vectors (and one shared step counter).  Values already in the
        optimiser's state -- i.e. just loaded from a checkpoint -- are copied into the flat vectors first."""
        with torch.no_grad():
            for p, off, n in self.param_slices:
                st = optimizer.state[p]
                m, v = self.exp_avg[off:off + n].view_as(p), self.exp_avg_sq[off:off + n].view_as(p)
                if "exp_avg" in st and st["exp_avg"].data_ptr() != m.data_ptr():
                    m.copy_(st["exp_avg"].to(self.device))
                    v.copy_(st["exp_avg_sq"].to(self.device))
                    self.step.copy_(torch.as_tensor(st["step"], dtype=torch.float32).to(self.device))
                st["exp_avg"], st["exp_avg_sq"], st["step"] = m, v, self.step

    def adam(self, n_samples, lr, loss_acc, adaptive, desired_kl, max_grad_norm, beta1, beta2, eps):
        """KL-adaptive learning rate + gradient-norm clip + Adam step on the flat vectors, one launch (nm_ppo_adam)."""
        a = PPOAdamArgs()
        a.n_params, a.n_samples, a.adaptive = self.flat.numel(), int(n_samples), 1 if adaptive else 0
        a.desired_kl, a.max_grad_norm = float(desired_kl or 0.0), float(max_grad_norm)
        a.beta1, a.beta2, a.eps = float(beta1), float(beta2), float(eps)
        a.params, a.grads, a.exp_avg, a.exp_avg_sq = self.flat.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.step, a.lr, a.sums = self.step.data_ptr(), lr.data_ptr(), self.out.data_ptr()
        a.loss_acc = loss_acc.data_ptr() if loss_acc is not None else None
        _lib.check(_lib.lib.nm_ppo_adam(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
