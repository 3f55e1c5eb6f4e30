"""Actor-critic MLP pair with rsl_rl v1.0.2's interface and ``state_dict`` keys
(``std``, ``actor.{0,2,4,...}.{weight,bias}``, ``critic.{0,2,4,...}.{weight,bias}``; consumed by the reference at
``play.py:65-72``).  Network sizes come from ``NightmareV3ConfigPPO.policy`` (``envs/nightmare_v3_config.py:105-109``)."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.distributions import Normal

_ACTIVATIONS = {
    "elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid,
}


def _activation(name: str) -> nn.Module:
    try:
        return _ACTIVATIONS[name]()
    except KeyError:
        raise ValueError(f"unknown activation '{name}' (one of {sorted(_ACTIVATIONS)})") from None


class _SplitKLinearFn(torch.autograd.Function):
    """y = x W^T + b whose weight gradient is a split-K batched GEMM.

    dW = grad_y^T x contracts over the whole mini-batch (81 920 rows at 4096 envs) into a 54x66 tile: cuBLAS runs that as
    ONE thread block walking K serially (~120 us).  Cutting the batch into chunks turns it into a batched GEMM that fills
    the GPU, followed by a tiny sum over the partial tiles."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gx = gy @ weight if ctx.needs_input_grad[0] else None
        s = x.shape[0]
        chunks = 0
        for c in (256, 160, 128, 100, 64, 50, 32):
            if s % c == 0 and s // c >= 64:
                chunks = c
                break
        if chunks:
            gw = torch.bmm(gy.view(chunks, s // chunks, -1).transpose(1, 2), x.view(chunks, s // chunks, -1)).sum(0)
            gb = gy.view(chunks, s // chunks, -1).sum(1).sum(0)
        else:
            gw = gy.t() @ x
            gb = gy.sum(0)
        return gx, gw, gb


class _Linear(nn.Linear):
    """nn.Linear (same parameters / state_dict keys) that routes large CUDA batches through the split-K backward."""

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and x.shape[0] >= 8192 and torch.is_grad_enabled() and self.weight.requires_grad:
            return _SplitKLinearFn.apply(x, self.weight, self.bias)
        return super().forward(x)


def _mlp(n_in: int, hidden, n_out: int, act: str) -> nn.Sequential:
    dims = [n_in, *hidden, n_out]
    layers = []
    for i in range(len(dims) - 1):
        layers.append(_Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(_activation(act))
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(256, 256, 256),
                 critic_hidden_dims=(256, 256, 256), activation="elu", init_noise_std=1.0, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        super().__init__()
        self.activation_name = activation
        self.actor = _mlp(num_actor_obs, list(actor_hidden_dims), num_actions, activation)
        self.critic = _mlp(num_critic_obs, list(critic_hidden_dims), 1, activation)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None

    # rsl_rl API -----------------------------------------------------------------------------------
    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, observations):
        mean = self.actor(observations)
        # validate_args=False: argument validation reads a reduction back on the host (a sync per call, and illegal
        # inside a CUDA-graph capture)
        self.distribution = Normal(mean, mean * 0.0 + self.std, validate_args=False)

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        return self.actor(observations)

    def evaluate(self, critic_observations, **kwargs):
        return self.critic(critic_observations)

    # helpers for the fused rollout kernel ----------------------------------------------------------
    def layer_dims(self):
        a = [m for m in self.actor if isinstance(m, nn.Linear)]
        c = [m for m in self.critic if isinstance(m, nn.Linear)]
        return ([a[0].in_features] + [m.out_features for m in a], [c[0].in_features] + [m.out_features for m in c])

    def flat_params(self):
        """(actor, critic) parameters flattened in module order: weight [out,in] then bias, layer by layer."""
        fa = torch.cat([p.detach().reshape(-1) for m in self.actor if isinstance(m, nn.Linear) for p in (m.weight, m.bias)])
        fc = torch.cat([p.detach().reshape(-1) for m in self.critic if isinstance(m, nn.Linear) for p in (m.weight, m.bias)])
        return fa.float().contiguous(), fc.float().contiguous()
