"""Rollout buffer with rsl_rl v1.0.2's layout (``[num_transitions_per_env, num_envs, ...]``), GAE(lambda) returns and
the shuffled mini-batch generator PPO.update consumes."""
from __future__ import annotations

import torch


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.clear()

        def clear(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device="cpu"):
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s: torch.zeros(T, N, *s, device=device)
        self.observations = z(*obs_shape)
        self.privileged_observations = z(*privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards, self.values, self.returns, self.advantages, self.actions_log_prob = z(1), z(1), z(1), z(1), z(1)
        self.actions, self.mu, self.sigma = z(*actions_shape), z(*actions_shape), z(*actions_shape)
        self.dones = torch.zeros(T, N, 1, device=device, dtype=torch.uint8)
        self.num_transitions_per_env, self.num_envs = T, N
        self.step = 0

    def add_transitions(self, t: "RolloutStorage.Transition"):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        i = self.step
        self.observations[i].copy_(t.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[i].copy_(t.critic_observations)
        self.actions[i].copy_(t.actions)
        self.rewards[i].copy_(t.rewards.view(-1, 1))
        self.dones[i].copy_(t.dones.view(-1, 1))
        self.values[i].copy_(t.values.view(-1, 1))
        self.actions_log_prob[i].copy_(t.actions_log_prob.view(-1, 1))
        self.mu[i].copy_(t.action_mean)
        self.sigma[i].copy_(t.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam, reduce_moments=None):
        """GAE(lambda); ``reduce_moments(sum, sumsq, count)`` lets a multi-process job normalise advantages with the
        moments of the GLOBAL batch (None: local batch, exactly rsl_rl's behaviour)."""
        if self.values.is_cuda and self.fused_gae:
            self._compute_returns_fused(last_values, gamma, lam, reduce_moments)
            return
        advantage = 0
        for step in reversed(range(self.num_transitions_per_env)):
            next_values = last_values if step == self.num_transitions_per_env - 1 else self.values[step + 1]
            not_terminal = 1.0 - self.dones[step].float()
            delta = self.rewards[step] + not_terminal * gamma * next_values - self.values[step]
            advantage = delta + not_terminal * gamma * lam * advantage
            self.returns[step] = advantage + self.values[step]
        adv = self.returns - self.values                    # written back IN PLACE: captured CUDA graphs keep reading this buffer
        if reduce_moments is None:
            self.advantages.copy_((adv - adv.mean()) / (adv.std() + 1e-8))
        else:
            a = adv.double()
            s, ss, n = reduce_moments(a.sum(), (a * a).sum(), torch.tensor(float(a.numel()), device=a.device, dtype=torch.float64))
            mean = s / n
            var = (ss - n * mean * mean) / (n - 1.0)          # unbiased, like torch.std
            self.advantages.copy_(((a - mean) / (var.clamp_min(0).sqrt() + 1e-8)).float())

    fused_gae = True

    def _compute_returns_fused(self, last_values, gamma, lam, reduce_moments):
        """Same recursion in one kernel (nm_gae): one thread per environment walks the T steps; the advantage moments come back
        in fp64 and the normalisation is three small tensor ops."""
        import ctypes

        from .. import _lib
        T, N = self.num_transitions_per_env, self.num_envs
        dev = self.values.device
        if getattr(self, "_moments", None) is None:
            self._moments = torch.zeros(2, device=dev, dtype=torch.float64)
        lv = last_values.detach().reshape(-1).float().contiguous()
        _lib.check(_lib.lib.nm_gae(T, N, self.rewards.data_ptr(), self.dones.data_ptr(), self.values.data_ptr(), lv.data_ptr(),
                                   float(gamma), float(lam), self.returns.data_ptr(), self.advantages.data_ptr(), self._moments.data_ptr(),
                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        s, ss = self._moments[0], self._moments[1]
        n = torch.tensor(float(T * N), device=dev, dtype=torch.float64)
        if reduce_moments is not None:
            s, ss, n = reduce_moments(s, ss, n)
        mean = s / n
        var = (ss - n * mean * mean) / (n - 1.0)              # unbiased, like torch.std
        self.advantages.sub_(mean.float()).div_(var.clamp_min(0).sqrt().float() + 1e-8)

    def get_statistics(self):
        done = self.dones.clone()
        done[-1] = 1
        flat = done.permute(1, 0, 2).reshape(-1, 1)
        idx = torch.cat((flat.new_tensor([-1], dtype=torch.int64), flat.nonzero(as_tuple=False)[:, 0]))
        lengths = idx[1:] - idx[:-1]
        return lengths.float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        batch = self.num_envs * self.num_transitions_per_env
        mb = batch // num_mini_batches
        indices = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        obs = self.observations.flatten(0, 1)
        cobs = self.privileged_observations.flatten(0, 1) if self.privileged_observations is not None else obs
        actions, values, returns = self.actions.flatten(0, 1), self.values.flatten(0, 1), self.returns.flatten(0, 1)
        logp, adv = self.actions_log_prob.flatten(0, 1), self.advantages.flatten(0, 1)
        mu, sigma = self.mu.flatten(0, 1), self.sigma.flatten(0, 1)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                b = indices[i * mb:(i + 1) * mb]
                yield (obs[b], cobs[b], actions[b], values[b], adv[b], returns[b], logp[b], mu[b], sigma[b], (None, None), None)
