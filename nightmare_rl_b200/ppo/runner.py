"""``OnPolicyRunner`` with rsl_rl v1.0.2's constructor / ``learn`` / ``save`` / ``load`` / ``get_inference_policy``
(the calls the reference makes at ``train.py:40,52,54`` and the checkpoint ``play.py:68-71`` reads).

Differences that matter on a GPU: episode statistics (the last-100-episodes reward / length means rsl_rl keeps in host
deques, which force a device->host sync on every step) are kept in device ring buffers and read once per iteration;
under ``torch.distributed`` only rank 0 logs and saves, and logged means are reduced over ranks."""
from __future__ import annotations

import os
import statistics
import time

import torch
import torch.distributed as dist

from .actor_critic import ActorCritic
from .ppo import PPO

_CLASSES = {"ActorCritic": ActorCritic, "PPO": PPO}


class _EpisodeRing:
    """Last ``cap`` finished episodes' (reward sum, length), updated without host sync."""

    def __init__(self, cap, device):
        self.cap = cap
        self.rew = torch.zeros(cap + 1, device=device)
        self.len = torch.zeros(cap + 1, device=device)
        self.count = torch.zeros(1, device=device, dtype=torch.int64)

    def push(self, done_mask, rew_sum, ep_len):
        d = done_mask.to(torch.int64)
        pos = (self.count + torch.cumsum(d, 0) - 1) % self.cap
        pos = torch.where(done_mask, pos, torch.full_like(pos, self.cap))       # non-finished envs write the spare slot
        self.rew.scatter_(0, pos, rew_sum)
        self.len.scatter_(0, pos, ep_len)
        self.count += d.sum()

    def means(self):
        n = int(min(int(self.count.item()), self.cap))
        if n == 0:
            return None, None, 0
        return float(self.rew[:n].mean()), float(self.len[:n].mean()), n


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cpu"):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.env = env
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        num_critic_obs = env.num_privileged_obs if env.num_privileged_obs is not None else env.num_obs
        ac_cls = _CLASSES[self.cfg["policy_class_name"]]
        actor_critic = ac_cls(env.num_obs, num_critic_obs, env.num_actions, **self.policy_cfg).to(self.device)
        alg_cls = _CLASSES[self.cfg["algorithm_class_name"]]
        extra = {}
        if alg_cls is PPO:
            extra = dict(seed=int(train_cfg.get("seed", 0)), env_offset=int(getattr(env, "env_offset", 0)))
        self.alg = alg_cls(actor_critic, device=self.device, **self.alg_cfg, **extra)
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg["save_interval"]
        self.alg.init_storage(env.num_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs], [env.num_actions])
        self.log_dir = log_dir if self.rank == 0 else None
        self.writer = None
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0.0, 0
        self.last_log = {}
        _, _ = self.env.reset()

    # ------------------------------------------------------------------ training loop
    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
            except Exception:                               # tensorboard missing: keep training, print only
                self.writer = None
        env, alg, dev = self.env, self.alg, self.device
        if init_at_random_ep_len:
            env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
        obs = env.get_observations()
        priv = env.get_privileged_observations()
        critic_obs = priv if priv is not None else obs
        obs, critic_obs = obs.to(dev), critic_obs.to(dev)
        alg.actor_critic.train()

        ep_infos = []
        ring = _EpisodeRing(100, dev)
        cur_rew = torch.zeros(env.num_envs, device=dev)
        cur_len = torch.zeros(env.num_envs, device=dev)
        # CUDA fast path: the transition store kernel also maintains these statistics (one launch per step)
        fused_stats = hasattr(alg, "attach_episode_stats") and alg.attach_episode_stats(cur_rew, cur_len, ring.rew[:ring.cap], ring.len[:ring.cap], ring.count)
        # pre-bound rollout: three C calls per env step, no tensors passed around on the host (CUDA env + fused policy only)
        ep_acc = torch.zeros(32, device=dev)
        fast = hasattr(alg, "prepare_fast_rollout") and self.cfg.get("fast_rollout", True) and alg.prepare_fast_rollout(env, ep_acc)
        first, last = self.current_learning_iteration, self.current_learning_iteration + int(num_learning_iterations)
        for it in range(first, last):
            start = time.time()
            with torch.inference_mode():
                for _ in range(self.num_steps_per_env if not fast else 0):
                    actions = alg.act(obs, critic_obs)
                    obs, priv, rewards, dones, infos = env.step(actions)
                    critic_obs = priv if priv is not None else obs
                    obs, critic_obs, rewards, dones = obs.to(dev), critic_obs.to(dev), rewards.to(dev), dones.to(dev)
                    alg.process_env_step(rewards, dones, infos)
                    if "episode" in infos:
                        ep_infos.append(infos["episode"])
                    if not fused_stats:
                        cur_rew += rewards
                        cur_len += 1
                        done_mask = dones > 0
                        ring.push(done_mask, cur_rew, cur_len)
                        cur_rew.masked_fill_(done_mask, 0.0)
                        cur_len.masked_fill_(done_mask, 0.0)
                if fast:
                    ep_acc.zero_()
                    for _ in range(self.num_steps_per_env):
                        alg.fast_rollout_step()
                    obs = critic_obs = env.get_observations()
                    n_ep = env._batch.ep_means.numel()
                    acc = ep_acc.tolist()
                    ep_infos.append({name: acc[i] / max(acc[n_ep], 1.0) for name, i in env._extras_keys})
                if dev.type == "cuda":
                    torch.cuda.synchronize(dev)
                stop = time.time()
                collection_time = stop - start
                start = stop
                alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = alg.update()
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            stop = time.time()
            learn_time = stop - start
            self._log(it, last, collection_time, learn_time, mean_value_loss, mean_surrogate_loss, ep_infos, ring)
            if self.log_dir is not None and it % self.save_interval == 0:
                self.save(os.path.join(self.log_dir, "model_{}.pt".format(it)))
            ep_infos.clear()
        self.current_learning_iteration += int(num_learning_iterations)
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, "model_{}.pt".format(self.current_learning_iteration)))

    def _log(self, it, last, collection_time, learn_time, v_loss, s_loss, ep_infos, ring, width=80, pad=35):
        steps = self.num_steps_per_env * self.env.num_envs * self.world
        self.tot_timesteps += steps
        iteration_time = collection_time + learn_time
        self.tot_time += iteration_time
        fps = int(steps / max(iteration_time, 1e-9))
        mean_rew, mean_len, n_ep = ring.means()
        mean_std = float(self.alg.actor_critic.std.detach().mean())
        ep_means = {}
        if ep_infos:
            for key in ep_infos[0]:
                vals = torch.stack([torch.as_tensor(e[key], device=self.device).float().reshape(()) for e in ep_infos])
                ep_means[key] = float(vals.mean())
        if self.world > 1:
            # every rank holds the statistics of its own env shard: reduce them (episode means weighted by the number of
            # finished episodes in each rank's ring, per-term episode means averaged) so that rank 0 logs the whole job
            keys = sorted(ep_means)
            loc = torch.tensor([(mean_rew or 0.0) * n_ep, (mean_len or 0.0) * n_ep, float(n_ep)] + [ep_means[k] for k in keys],
                               device=self.device, dtype=torch.float64)
            dist.all_reduce(loc)
            n_all = float(loc[2])
            if n_all > 0:
                mean_rew, mean_len, n_ep = float(loc[0]) / n_all, float(loc[1]) / n_all, int(n_all)
            for i, k in enumerate(keys):
                ep_means[k] = float(loc[3 + i]) / self.world
        self.last_log = dict(iteration=it, fps=fps, collection_time=collection_time, learn_time=learn_time, value_loss=v_loss,
                             surrogate_loss=s_loss, mean_reward=mean_rew, mean_episode_length=mean_len, mean_noise_std=mean_std,
                             learning_rate=self.alg.learning_rate, episode=ep_means, total_timesteps=self.tot_timesteps)
        if self.rank != 0:
            return
        if self.writer is not None:
            w = self.writer
            for key, v in ep_means.items():
                w.add_scalar("Episode/" + key, v, it)
            w.add_scalar("Loss/value_function", v_loss, it)
            w.add_scalar("Loss/surrogate", s_loss, it)
            w.add_scalar("Loss/learning_rate", self.alg.learning_rate, it)
            w.add_scalar("Policy/mean_noise_std", mean_std, it)
            w.add_scalar("Perf/total_fps", fps, it)
            w.add_scalar("Perf/collection time", collection_time, it)
            w.add_scalar("Perf/learning_time", learn_time, it)
            if n_ep > 0:
                w.add_scalar("Train/mean_reward", mean_rew, it)
                w.add_scalar("Train/mean_episode_length", mean_len, it)
                w.add_scalar("Train/mean_reward/time", mean_rew, self.tot_time)
                w.add_scalar("Train/mean_episode_length/time", mean_len, self.tot_time)
        if self.log_dir is None and not self.cfg.get("verbose", False):
            return
        head = f" \033[1m Learning iteration {it}/{last} \033[0m "
        lines = ["#" * width, head.center(width, " "), "",
                 f"{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {collection_time:.3f}s, learning {learn_time:.3f}s)",
                 f"{'Value function loss:':>{pad}} {v_loss:.4f}", f"{'Surrogate loss:':>{pad}} {s_loss:.4f}",
                 f"{'Mean action noise std:':>{pad}} {mean_std:.2f}"]
        if n_ep > 0:
            lines += [f"{'Mean reward:':>{pad}} {mean_rew:.2f}", f"{'Mean episode length:':>{pad}} {mean_len:.2f}"]
        lines += [f"{'Mean episode ' + k + ':':>{pad}} {v:.4f}" for k, v in ep_means.items()]
        eta = self.tot_time / (it + 1 - (self.current_learning_iteration)) * (last - it - 1) if last < 10 ** 8 else float("nan")
        lines += ["-" * width, f"{'Total timesteps:':>{pad}} {self.tot_timesteps}", f"{'Iteration time:':>{pad}} {iteration_time:.2f}s",
                  f"{'Total time:':>{pad}} {self.tot_time:.2f}s", f"{'ETA:':>{pad}} {eta:.1f}s"]
        print("\n".join(lines))

    # ------------------------------------------------------------------ checkpoints (format read by play.py:68-71)
    def save(self, path, infos=None):
        if self.rank != 0:
            return
        opt = self.alg.optimizer.state_dict()
        for g in opt["param_groups"]:                       # plain-float learning rate, as rsl_rl's checkpoints have it
            if torch.is_tensor(g.get("lr")):
                g["lr"] = float(g["lr"])
        torch.save({"model_state_dict": self.alg.actor_critic.state_dict(), "optimizer_state_dict": opt,
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        loaded = torch.load(path, map_location=self.device, weights_only=False)
        self.alg.actor_critic.load_state_dict(loaded["model_state_dict"])
        if load_optimizer:
            self.alg.optimizer.load_state_dict(loaded["optimizer_state_dict"])
            if hasattr(self.alg, "optimizer_reloaded"):
                self.alg.optimizer_reloaded()
        self.current_learning_iteration = loaded["iter"]
        if hasattr(self.alg, "weights_changed"):
            self.alg.weights_changed()
        return loaded["infos"]

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference
