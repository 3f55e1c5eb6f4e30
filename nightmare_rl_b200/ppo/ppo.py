"""Proximal Policy Optimisation with rsl_rl v1.0.2's interface and update rule (hyper-parameters:
``NightmareV3ConfigPPO.algorithm``, reference ``envs/nightmare_v3_config.py:117-131``): clipped surrogate, clipped value
loss, entropy bonus, gradient-norm clipping, KL-adaptive learning rate, ``num_learning_epochs x num_mini_batches`` shuffled
mini-batches, time-out bootstrapping of the reward.

Multi-GPU: when ``torch.distributed`` is initialised each rank owns a shard of the environments; the gradients of every
optimiser step are averaged with ONE all-reduce of a flat buffer (15 043 parameters here), the mean KL is all-reduced so
every rank takes the same learning-rate decision, and advantages are normalised with global moments."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim

from .storage import RolloutStorage


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class PPO:
    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cpu", fused_rollout=None, seed=0, env_offset=0, graph_update=None):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.desired_kl, self.schedule, self.learning_rate = desired_kl, schedule, learning_rate
        self.actor_critic = actor_critic.to(self.device)
        self.storage = None
        self.optimizer = optim.Adam(self.actor_critic.parameters(), lr=learning_rate)
        self.transition = RolloutStorage.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef, self.gamma, self.lam = value_loss_coef, entropy_coef, gamma, lam
        self.max_grad_norm, self.use_clipped_value_loss = max_grad_norm, use_clipped_value_loss
        self.world = _world()
        if self.world > 1:                                  # identical initial weights on every rank
            for p in self.actor_critic.parameters():
                dist.broadcast(p.data, src=0)
        # fused tensor-core rollout forward (CUDA only; the PyTorch modules remain the differentiable path for update())
        if fused_rollout is None:
            fused_rollout = self.device.type == "cuda"
        self.fused = None
        self._seed, self._env_offset = seed, env_offset
        if fused_rollout:
            from .policy_kernel import FusedPolicy
            self.fused = FusedPolicy(self.actor_critic, self.device, seed=seed, env_offset=env_offset)
        self._weights_dirty = False
        self._act_calls = 0
        self._flat_grad = None
        self._stats = None
        self._in_place = False
        self.tf32_backward = True
        self.fused_head = self.device.type == "cuda"
        # whole mini-batch gradient (gather + both MLPs forward/backward + loss head) in one kernel; needs ELU networks that
        # fit one SM's shared memory, otherwise the autograd path (with the fused head) stays
        self.fused_grad = None
        self._want_fused_grad = self.device.type == "cuda" and os.environ.get("NM_PPO_FUSED_GRAD", "1") != "0"
        # CUDA-graph replay of the mini-batch update (CUDA runs; NCCL collectives are captured too): ~200 tiny kernels per
        # mini-batch and otherwise bound by PyTorch's per-op launch overhead.  The optimiser then keeps its learning rate
        # in a device tensor and the KL-adaptive schedule runs on the device as well (same rule, no host read-back).
        if graph_update is None:
            graph_update = self.device.type == "cuda"
        self.graph_update = bool(graph_update) and self.device.type == "cuda"
        self._graph = None
        if self.graph_update:
            self._lr_t = torch.tensor(float(learning_rate), device=self.device)
            if self._want_fused_grad:
                try:
                    from .policy_kernel import FusedPPOGrad
                    self.fused_grad = FusedPPOGrad(self.actor_critic, self.device)
                except Exception:                              # unsupported network: autograd path
                    self.fused_grad = None
            self.optimizer = optim.Adam(self.actor_critic.parameters(), lr=self._lr_t, capturable=True, foreach=True)
            if self.fused_grad is not None:
                # the update itself is nm_ppo_adam on the flat vectors; the torch optimiser object stays as the holder of
                # hyper-parameters and of the (re-seated) state, so checkpoints keep rsl_rl's optimizer_state_dict layout
                g0 = self.optimizer.param_groups[0]
                if g0.get("amsgrad") or g0.get("weight_decay") or g0.get("maximize"):
                    self.fused_grad = None
                else:
                    self.fused_grad.seat_optimizer_state(self.optimizer)

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def weights_changed(self):
        """Call after loading a checkpoint: the packed copy the fused kernel reads is refreshed before the next act()."""
        self._weights_dirty = True

    def release_graph(self):
        """Drop the captured update graph (it holds NCCL work when ranks > 1; release it before the process group goes)."""
        self._graph = None
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def optimizer_reloaded(self):
        """After ``optimizer.load_state_dict``: re-attach the device-resident learning rate and drop the captured graph
        (the optimiser's state tensors have new addresses)."""
        lr = self.optimizer.param_groups[0]["lr"]
        self.learning_rate = float(lr)
        if self.graph_update:
            self._lr_t.fill_(self.learning_rate)
            for g in self.optimizer.param_groups:
                g["lr"] = self._lr_t
                g["capturable"] = True
            for stt in self.optimizer.state.values():
                if "step" in stt and torch.is_tensor(stt["step"]):
                    stt["step"] = stt["step"].to(self.device)
            if self.fused_grad is not None:
                self.fused_grad.seat_optimizer_state(self.optimizer)
            self._graph = None

    # ------------------------------------------------------------------ rollout
    def act(self, obs, critic_obs):
        t = self.transition
        self._act_calls += 1
        t_ptr = obs.data_ptr()
        if self.fused is not None:
            if self._weights_dirty:
                self.fused.load(self.actor_critic)
                self._weights_dirty = False
            st = self.storage
            out = None
            self._in_place = st is not None and st.step < st.num_transitions_per_env and obs.shape[0] == st.num_envs
            oc = sg = None
            if self._in_place:                               # the kernel writes straight into row `step` of the rollout buffer,
                i = st.step                                  # including the observations it saw: a GPU env returns VIEWS of buffers
                out = (st.actions[i], st.mu[i], st.values[i].view(-1), st.actions_log_prob[i].view(-1))   # it overwrites in step()
                oc, sg = st.observations[i], st.sigma[i]
            actions, mean, value, logp = self.fused.act(obs, self._act_calls, out=out, obs_copy=oc, sigma_out=sg)
            t.actions, t.values, t.actions_log_prob, t.action_mean = actions, value.unsqueeze(1), logp, mean
            t.action_sigma = self.fused.std.unsqueeze(0).expand_as(mean)
            if self._in_place:                               # the transition refers to the snapshot, not to the env's buffer
                critic_obs = oc if critic_obs is None or critic_obs.data_ptr() == t_ptr else critic_obs.clone()
                obs = oc
            else:
                obs = obs.clone()
                critic_obs = obs if critic_obs is None or critic_obs.data_ptr() == t_ptr else critic_obs.clone()
        else:
            if obs.is_cuda:                                  # snapshot: the env may overwrite the buffer `obs` is a view of
                obs = obs.clone()
                critic_obs = obs if critic_obs is None or critic_obs.data_ptr() == t_ptr else critic_obs.clone()
            t.actions = self.actor_critic.act(obs).detach()
            t.values = self.actor_critic.evaluate(critic_obs).detach()
            t.actions_log_prob = self.actor_critic.get_actions_log_prob(t.actions).detach()
            t.action_mean = self.actor_critic.action_mean.detach()
            t.action_sigma = self.actor_critic.action_std.detach()
        t.observations, t.critic_observations = obs, critic_obs
        return t.actions

    def attach_episode_stats(self, cur_rew, cur_len, ring_rew, ring_len, ring_count):
        """Runner hook: running reward / length per env and the ring of finished episodes are then updated by the same
        launch that stores the transition (CUDA fast path only; returns whether it is active)."""
        self._stats = (cur_rew, cur_len, ring_rew, ring_len, ring_count)
        return self.fused is not None

    def _store_fused(self, rewards, dones, infos):
        from .policy_kernel import RolloutSlot, rollout_store
        t, st = self.transition, self.storage
        i = st.step
        if i >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        obs = t.observations
        tout = infos.get("time_outs") if isinstance(infos, dict) else None
        ok = (obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and rewards.is_cuda and rewards.dtype == torch.float32
              and rewards.is_contiguous() and dones.is_cuda and dones.dtype == torch.int64 and dones.is_contiguous()
              and (tout is None or (tout.is_cuda and tout.dtype == torch.float32 and tout.is_contiguous()))
              and st.privileged_observations is None)
        if not ok:
            return False
        s = RolloutSlot()
        s.n, s.obs_dim, s.act_dim = st.num_envs, obs.shape[1], t.actions.shape[1]
        s.gamma = float(self.gamma)
        if not self._in_place:
            s.obs, s.std = obs.data_ptr(), self.fused.std.data_ptr()   # (callers that bypassed the in-place rows must pass a snapshot)
            s.actions, s.mean = t.actions.contiguous().data_ptr(), t.action_mean.contiguous().data_ptr()
            s.value, s.logp = t.values.contiguous().data_ptr(), t.actions_log_prob.contiguous().data_ptr()
        s.rew, s.done = rewards.data_ptr(), dones.data_ptr()
        s.time_outs = tout.data_ptr() if tout is not None else None
        s.s_obs, s.s_actions, s.s_mu, s.s_sigma = st.observations[i].data_ptr(), st.actions[i].data_ptr(), st.mu[i].data_ptr(), st.sigma[i].data_ptr()
        s.s_values, s.s_logp = st.values[i].data_ptr(), st.actions_log_prob[i].data_ptr()
        s.s_rewards, s.s_dones = st.rewards[i].data_ptr(), st.dones[i].data_ptr()
        if self._stats is not None:
            cr, cl, rr, rl, rc = self._stats
            s.cur_rew, s.cur_len, s.ring_rew, s.ring_len, s.ring_count = cr.data_ptr(), cl.data_ptr(), rr.data_ptr(), rl.data_ptr(), rc.data_ptr()
            s.ring_cap = rr.numel()
        rollout_store(s, self.device)
        st.step += 1
        t.clear()
        return True

    # ------------------------------------------------------------------ pre-bound rollout (host cost: three C calls per step)
    def prepare_fast_rollout(self, env, ep_acc=None):
        """Bind, once, every pointer a rollout step needs (env output buffers, row t of the rollout buffer, statistics)
        so that a step is nm_policy_act -> nm_step -> nm_rollout_store with pre-built argument blocks.  Returns False
        when the combination is not eligible (no fused kernel, foreign env class, privileged observations...)."""
        import ctypes
        from .policy_kernel import RolloutSlot
        b = getattr(env, "_batch", None)
        st = self.storage
        if (self.fused is None or b is None or st is None or b.device != self.device or env.num_envs != st.num_envs
                or getattr(env, "_copy", False) or not env.cfg.env.send_timeouts or env.get_privileged_observations() is not None):
            return False
        # the env has no privileged observations (the critic sees the actor's): the separate critic-observation rows the
        # runner allocated (num_privileged_obs == num_obs, reference quirk) would only ever hold a copy of `observations`
        st.privileged_observations = None
        f = self.fused
        self._fast = []
        for i in range(st.num_transitions_per_env):
            s = RolloutSlot()
            s.n, s.obs_dim, s.act_dim, s.gamma = st.num_envs, b.obs.shape[1], st.actions.shape[2], float(self.gamma)
            s.rew, s.done, s.time_outs = b.rew.data_ptr(), b.done.data_ptr(), b.time_outs_latched.data_ptr()      # obs/std rows: written by the policy launch
            s.s_obs, s.s_actions, s.s_mu, s.s_sigma = st.observations[i].data_ptr(), st.actions[i].data_ptr(), st.mu[i].data_ptr(), st.sigma[i].data_ptr()
            s.s_values, s.s_logp, s.s_rewards, s.s_dones = st.values[i].data_ptr(), st.actions_log_prob[i].data_ptr(), st.rewards[i].data_ptr(), st.dones[i].data_ptr()
            if self._stats is not None:
                cr, cl, rr, rl, rc = self._stats
                s.cur_rew, s.cur_len, s.ring_rew, s.ring_len, s.ring_count = cr.data_ptr(), cl.data_ptr(), rr.data_ptr(), rl.data_ptr(), rc.data_ptr()
                s.ring_cap = rr.numel()
            if ep_acc is not None:
                s.ep_means, s.ep_acc, s.n_ep = b.ep_means.data_ptr(), ep_acc.data_ptr(), b.ep_means.numel()
            self._fast.append((s, st.actions[i]))
        self._fast_env = env
        return True

    def fast_rollout_step(self):
        """One env step of the rollout through the pre-bound pointers (obs are read from / written to the env's own
        persistent buffers, so nothing is passed around on the host)."""
        import ctypes
        from .. import _lib
        env, b, f, st = self._fast_env, self._fast_env._batch, self.fused, self.storage
        i = st.step
        if i >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        if self._weights_dirty:
            f.load(self.actor_critic)
            self._weights_dirty = False
        slot, act_row = self._fast[i]
        self._act_calls += 1
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        # the policy launch also writes the rollout buffer's copy of the observations and the std row (the env is about
        # to overwrite its observation buffer); the post-step store then only adds rewards / dones / statistics
        f._launches += 1
        _lib.check(f._fn[3](f._h, b.obs.data_ptr(), slot.obs_dim, slot.n, ctypes.c_uint64(f.seed), ctypes.c_int64(self._act_calls),
                                                ctypes.c_int64(f.env_offset), 0, slot.s_actions, slot.s_mu, slot.s_values, slot.s_logp,
                                                slot.s_obs, slot.s_sigma, stream))
        env.fast_step(act_row)
        _lib.check(_lib.lib.nm_rollout_store(ctypes.byref(slot), stream))
        st.step += 1

    def process_env_step(self, rewards, dones, infos):
        if self.fused is not None and self._store_fused(rewards, dones, infos):
            self.actor_critic.reset(dones)
            return
        t = self.transition
        t.rewards = rewards.clone()
        t.dones = dones
        if "time_outs" in infos:                            # bootstrap the value of episodes cut by the time limit
            t.rewards += self.gamma * torch.squeeze(t.values * infos["time_outs"].unsqueeze(1).to(self.device), 1)
        self.storage.add_transitions(t)
        t.clear()
        if self._stats is not None:                         # statistics the runner delegated to this class (slow path)
            cr, cl, rr, rl, rc = self._stats
            cr += rewards
            cl += 1
            dm = dones > 0
            d = dm.to(torch.int64)
            pos = (rc + torch.cumsum(d, 0) - 1) % rr.numel()
            keep = torch.nonzero(dm).flatten()
            rr[pos[keep]] = cr[keep]
            rl[pos[keep]] = cl[keep]
            rc += d.sum()
            cr.masked_fill_(dm, 0.0)
            cl.masked_fill_(dm, 0.0)
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        last_values = self.actor_critic.evaluate(last_critic_obs).detach()
        self.storage.compute_returns(last_values, self.gamma, self.lam, self._reduce_moments if self.world > 1 else None)

    # ------------------------------------------------------------------ collectives
    @staticmethod
    def _reduce_moments(s, ss, n):
        buf = torch.stack([s, ss, n])
        dist.all_reduce(buf)
        return buf[0], buf[1], buf[2]

    def _allreduce_grads(self):
        params = [p for p in self.actor_critic.parameters() if p.grad is not None]
        total = sum(p.numel() for p in params)
        if self._flat_grad is None or self._flat_grad.numel() != total:
            self._flat_grad = torch.empty(total, device=self.device)
        torch.cat([p.grad.reshape(-1) for p in params], out=self._flat_grad)
        dist.all_reduce(self._flat_grad)
        self._flat_grad.div_(self.world)
        off = 0
        for p in params:
            p.grad.copy_(self._flat_grad[off:off + p.numel()].view_as(p.grad))
            off += p.numel()

    def _backward(self, loss):
        """Backward pass with TF32 tensor-core GEMMs on CUDA.  The forward pass stays fp32 (the probability ratio against the
        rollout's log-probs must be exact); the weight-gradient GEMMs contract over the whole mini-batch (K = 81 920 at 4096
        envs) and are what cuBLAS otherwise runs as CUDA-core SIMT kernels — 26 of the 48 ms of an update."""
        if self.device.type != "cuda" or not self.tf32_backward:
            loss.backward()
            return
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            loss.backward()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    # ------------------------------------------------------------------ update
    def update(self):
        if self.graph_update:
            return self._update_graphed()
        return self._update_eager()

    # one mini-batch of the update rule on explicit tensors (shared by the eager and the graph-captured path)
    def _minibatch_loss(self, obs_b, cobs_b, act_b, tgt_val_b, adv_b, ret_b, old_logp_b, old_mu_b, old_sigma_b):
        ac = self.actor_critic
        if self.fused_head and obs_b.is_cuda:
            # networks through autograd + cuBLAS, everything after them (Normal log-prob, ratio, clipped surrogate, clipped value
            # loss, entropy, KL) and its gradients in one kernel
            from .policy_kernel import FusedPPOHead
            mu_b = ac.actor(obs_b)
            value_b = ac.evaluate(cobs_b)
            loss, value_loss, surrogate_loss, kl_mean = FusedPPOHead.apply(
                mu_b, value_b, ac.std, act_b, old_logp_b, old_mu_b, old_sigma_b, adv_b, ret_b, tgt_val_b,
                self.clip_param, self.value_loss_coef, self.entropy_coef, self.use_clipped_value_loss)
            return loss, value_loss.detach(), surrogate_loss.detach(), kl_mean.detach()
        ac = self.actor_critic
        ac.update_distribution(obs_b)                      # rsl_rl calls act() here and discards the sample
        logp_b = ac.get_actions_log_prob(act_b)
        value_b = ac.evaluate(cobs_b)
        mu_b, sigma_b, entropy_b = ac.action_mean, ac.action_std, ac.entropy
        with torch.no_grad():
            kl = torch.sum(torch.log(sigma_b / old_sigma_b + 1.0e-5)
                           + (torch.square(old_sigma_b) + torch.square(old_mu_b - mu_b)) / (2.0 * torch.square(sigma_b)) - 0.5, axis=-1)
            kl_mean = torch.mean(kl)
        ratio = torch.exp(logp_b - torch.squeeze(old_logp_b))
        adv = torch.squeeze(adv_b)
        surrogate = -adv * ratio
        surrogate_clipped = -adv * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
        surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
        if self.use_clipped_value_loss:
            value_clipped = tgt_val_b + (value_b - tgt_val_b).clamp(-self.clip_param, self.clip_param)
            value_loss = torch.max((value_b - ret_b).pow(2), (value_clipped - ret_b).pow(2)).mean()
        else:
            value_loss = (ret_b - value_b).pow(2).mean()
        loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_b.mean()
        return loss, value_loss.detach(), surrogate_loss.detach(), kl_mean

    def _graph_step(self):
        """Body that is captured once and replayed per mini-batch: gather by the static index tensor, loss, backward,
        device-side KL-adaptive learning rate, gradient clipping, Adam."""
        st, b = self.storage, self._g_idx
        obs = st.observations.flatten(0, 1)
        cobs = st.privileged_observations.flatten(0, 1) if st.privileged_observations is not None else obs
        fg = self.fused_grad
        if fg is not None:
            # one kernel: gather + forward + loss head + backward of both networks, gradient left in fg.flat_grad (= every p.grad)
            n = b.numel()
            out = fg(n, b, obs, cobs, st.actions.flatten(0, 1), st.actions_log_prob.flatten(0, 1), st.mu.flatten(0, 1), st.sigma.flatten(0, 1),
                     st.advantages.flatten(0, 1), st.returns.flatten(0, 1), st.values.flatten(0, 1),
                     self.clip_param, self.value_loss_coef, self.entropy_coef, self.use_clipped_value_loss)
            if self.world > 1:                               # ONE NCCL all-reduce (captured with the rest) averages the flat gradient
                dist.all_reduce(fg.ext)                      # and the loss / KL sums, so every rank takes the same learning-rate decision
                fg.ext.div_(self.world)
        else:
            loss, v_loss, s_loss, kl_mean = self._minibatch_loss(
                obs[b], cobs[b], st.actions.flatten(0, 1)[b], st.values.flatten(0, 1)[b], st.advantages.flatten(0, 1)[b],
                st.returns.flatten(0, 1)[b], st.actions_log_prob.flatten(0, 1)[b], st.mu.flatten(0, 1)[b], st.sigma.flatten(0, 1)[b])
            if self.world > 1:                               # every rank takes the same learning-rate decision
                kl_mean = kl_mean.clone()
                dist.all_reduce(kl_mean)
                kl_mean = kl_mean / self.world
        if fg is not None:
            # KL-adaptive learning rate, clip_grad_norm_ and the Adam step in one launch on the flat vectors
            g0 = self.optimizer.param_groups[0]
            fg.adam(n, self._lr_t, self._g_losses, self.desired_kl is not None and self.schedule == "adaptive", self.desired_kl,
                    self.max_grad_norm, g0["betas"][0], g0["betas"][1], g0["eps"])
            return
        if self.desired_kl is not None and self.schedule == "adaptive":
            lr = self._lr_t
            down = torch.clamp(lr / 1.5, min=1e-5)
            up = torch.clamp(lr * 1.5, max=1e-2)
            new_lr = torch.where(kl_mean > self.desired_kl * 2.0, down,
                                 torch.where((kl_mean < self.desired_kl / 2.0) & (kl_mean > 0.0), up, lr))
            lr.copy_(new_lr)
        self.optimizer.zero_grad(set_to_none=False)
        self._backward(loss)
        if self.world > 1:
            self._allreduce_grads()
        nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm, foreach=True)
        self.optimizer.step()
        self._g_vloss += v_loss
        self._g_sloss += s_loss

    def _update_graphed(self):
        st = self.storage
        batch = st.num_envs * st.num_transitions_per_env
        mb = batch // self.num_mini_batches
        if self._graph is None:
            self._g_idx = torch.zeros(mb, dtype=torch.int64, device=self.device)
            self._g_losses = torch.zeros(2, device=self.device)           # [value loss, surrogate loss] accumulated over the mini-batches
            self._g_vloss, self._g_sloss = self._g_losses[0], self._g_losses[1]
            # gradients must exist (and keep their addresses) before capture
            for p in self.actor_critic.parameters():
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            snap = [p.detach().clone() for p in self.actor_critic.parameters()]
            # optimiser state that already exists (moments / step reloaded from a checkpoint by `train.py -r`, or carried over
            # a release_graph()) must survive the warm-up steps: snapshot it, restore it afterwards
            opt_snap = {id(v): v.detach().clone() for stt in self.optimizer.state.values() for v in stt.values() if torch.is_tensor(v)}
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):               # warm-up on a side stream (allocator / cuBLAS / optimiser state)
                self._g_idx.copy_(torch.arange(mb, device=self.device))     # (does not touch the global RNG stream)
                lr_keep = self._lr_t.clone()
                for _ in range(2):
                    self._graph_step()
                # undo the warm-up's effect on the parameters, the schedule and Adam's state (state created by the warm-up
                # itself restarts from zero, state that existed before continues from its values)
                with torch.no_grad():
                    for p, q in zip(self.actor_critic.parameters(), snap):
                        p.copy_(q)
                    self._lr_t.copy_(lr_keep)
                    for stt in self.optimizer.state.values():
                        for k, v in stt.items():
                            if torch.is_tensor(v):
                                if id(v) in opt_snap:
                                    v.copy_(opt_snap[id(v)])
                                else:
                                    v.zero_()
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._graph_step()
            # the capture itself does not execute; state is as after the restore above
        self._g_vloss.zero_()
        self._g_sloss.zero_()
        indices = torch.randperm(self.num_mini_batches * mb, device=self.device)
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._g_idx.copy_(indices[i * mb:(i + 1) * mb])
                self._graph.replay()
        n_upd = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        self._weights_dirty = True
        out = torch.stack([self._g_vloss / n_upd, self._g_sloss / n_upd, self._lr_t]).tolist()      # one host read per iteration
        self.learning_rate = out[2]
        return out[0], out[1]

    def _update_eager(self):
        mean_value_loss = torch.zeros((), device=self.device)
        mean_surrogate_loss = torch.zeros((), device=self.device)
        ac = self.actor_critic
        gen = self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        for (obs_b, cobs_b, act_b, tgt_val_b, adv_b, ret_b, old_logp_b, old_mu_b, old_sigma_b, _hid, _mask) in gen:
            ac.act(obs_b)
            logp_b = ac.get_actions_log_prob(act_b)
            value_b = ac.evaluate(cobs_b)
            mu_b, sigma_b, entropy_b = ac.action_mean, ac.action_std, ac.entropy

            if self.desired_kl is not None and self.schedule == "adaptive":
                with torch.inference_mode():
                    kl = torch.sum(torch.log(sigma_b / old_sigma_b + 1.0e-5)
                                   + (torch.square(old_sigma_b) + torch.square(old_mu_b - mu_b)) / (2.0 * torch.square(sigma_b)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    if self.world > 1:
                        dist.all_reduce(kl_mean)
                        kl_mean /= self.world
                    kl_val = float(kl_mean)                # one host read per mini-batch, as in rsl_rl
                    if kl_val > self.desired_kl * 2.0:
                        self.learning_rate = max(1e-5, self.learning_rate / 1.5)
                    elif 0.0 < kl_val < self.desired_kl / 2.0:
                        self.learning_rate = min(1e-2, self.learning_rate * 1.5)
                    for g in self.optimizer.param_groups:
                        g["lr"] = self.learning_rate

            ratio = torch.exp(logp_b - torch.squeeze(old_logp_b))
            adv = torch.squeeze(adv_b)
            surrogate = -adv * ratio
            surrogate_clipped = -adv * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
            surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
            if self.use_clipped_value_loss:
                value_clipped = tgt_val_b + (value_b - tgt_val_b).clamp(-self.clip_param, self.clip_param)
                value_loss = torch.max((value_b - ret_b).pow(2), (value_clipped - ret_b).pow(2)).mean()
            else:
                value_loss = (ret_b - value_b).pow(2).mean()
            loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_b.mean()

            self.optimizer.zero_grad()
            self._backward(loss)
            if self.world > 1:
                self._allreduce_grads()
            nn.utils.clip_grad_norm_(ac.parameters(), self.max_grad_norm)
            self.optimizer.step()
            mean_value_loss += value_loss.detach()
            mean_surrogate_loss += surrogate_loss.detach()

        n_upd = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        self._weights_dirty = True
        return float(mean_value_loss / n_upd), float(mean_surrogate_loss / n_upd)
