"""PPO runner / algorithm / storage (nightmare_rl_b200.ppo ≙ rsl_rl v1.0.2, the runner the reference imports at
train.py:1 and play.py:11-12) on CPU: shapes, GAE, update rule, checkpoint format, rsl_rl import shim, and the
world_size-2 gradient all-reduce (gloo)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


class FakeEnv:
    """Duck-typed stand-in with the attribute surface rsl_rl touches, CPU tensors."""

    def __init__(self, n=32, num_obs=66, num_actions=18, seed=0):
        self.num_envs, self.num_obs, self.num_privileged_obs, self.num_actions = n, num_obs, num_obs, num_actions
        self.max_episode_length = np.float64(50.0)
        self.episode_length_buf = torch.zeros(n, dtype=torch.int64)
        self.g = torch.Generator().manual_seed(seed)
        self.obs = torch.zeros(n, num_obs)
        self.rebinds = 0

    def reset(self):
        self.obs = torch.randn(self.num_envs, self.num_obs, generator=self.g)
        return self.obs, None

    def get_observations(self):
        return self.obs

    def get_privileged_observations(self):
        return None

    def step(self, actions):
        assert actions.shape == (self.num_envs, self.num_actions)
        self.episode_length_buf += 1
        time_out = self.episode_length_buf > int(self.max_episode_length)
        fell = torch.rand(self.num_envs, generator=self.g) < 0.02
        done = (time_out | fell)
        rew = -(actions ** 2).mean(dim=1) * 0.01 + 0.1 * self.obs[:, 0]
        self.episode_length_buf[done] = 0
        self.obs = torch.randn(self.num_envs, self.num_obs, generator=self.g)
        extras = {"time_outs": time_out.float(), "episode": {"rew_x": torch.tensor(0.5)}}
        return self.obs, None, rew, done.to(torch.int64), extras


def _train_cfg():
    from nightmare_rl_b200.envs.helpers import class_to_dict
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3ConfigPPO
    tc = NightmareV3ConfigPPO()
    tc.runner.num_steps_per_env = 8
    tc.runner.save_interval = 1
    return class_to_dict(tc)


def test_actor_critic_matches_reference_network():
    from nightmare_rl_b200.ppo import ActorCritic
    ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30], activation="elu", init_noise_std=1.0)
    keys = set(ac.state_dict().keys())
    want = {"std"} | {f"{net}.{i}.{p}" for net in ("actor", "critic") for i in (0, 2, 4, 6) for p in ("weight", "bias")}
    assert keys == want                                             # layout play.py:71 load_state_dict expects
    assert sum(p.numel() for p in ac.parameters()) == 15043
    obs = torch.randn(5, 66)
    a = ac.act(obs)
    assert a.shape == (5, 18) and ac.get_actions_log_prob(a).shape == (5,) and ac.evaluate(obs).shape == (5, 1)
    assert torch.allclose(ac.action_std, torch.ones(5, 18)) and ac.act_inference(obs).shape == (5, 18)
    fa, fc = ac.flat_params()
    assert fa.numel() == 66 * 54 + 54 + 54 * 42 + 42 + 42 * 30 + 30 + 30 * 18 + 18 and fc.numel() == 66 * 54 + 54 + 54 * 42 + 42 + 42 * 30 + 30 + 30 + 1


def test_gae_against_plain_loop():
    from nightmare_rl_b200.ppo import RolloutStorage
    T, N, gamma, lam = 12, 7, 0.99, 0.95
    st = RolloutStorage(N, T, [3], [None], [2])
    g = torch.Generator().manual_seed(0)
    st.rewards.copy_(torch.randn(T, N, 1, generator=g))
    st.values.copy_(torch.randn(T, N, 1, generator=g))
    st.dones.copy_((torch.rand(T, N, 1, generator=g) < 0.2).to(torch.uint8))
    last = torch.randn(N, 1, generator=g)
    st.compute_returns(last, gamma, lam)
    r, v, d = st.rewards.numpy()[..., 0], st.values.numpy()[..., 0], st.dones.numpy()[..., 0].astype(float)
    ret = np.zeros((T, N))
    for n in range(N):                                             # textbook GAE(lambda), one env at a time
        adv = 0.0
        for t in reversed(range(T)):
            nv = last.numpy()[n, 0] if t == T - 1 else v[t + 1, n]
            nt = 1.0 - d[t, n]
            delta = r[t, n] + nt * gamma * nv - v[t, n]
            adv = delta + nt * gamma * lam * adv
            ret[t, n] = adv + v[t, n]
    assert np.allclose(st.returns.numpy()[..., 0], ret, atol=1e-5)
    a = ret - v
    assert np.allclose(st.advantages.numpy()[..., 0], (a - a.mean()) / (a.std(ddof=1) + 1e-8), atol=1e-4)
    batches = list(st.mini_batch_generator(4, 5))
    assert len(batches) == 20 and batches[0][0].shape == (T * N // 4, 3) and batches[0][2].shape == (T * N // 4, 2)


def test_runner_learn_save_load(tmp_path):
    from rsl_rl.runners import OnPolicyRunner                       # the import the reference's train.py:1 performs
    env = FakeEnv()
    log_dir = str(tmp_path / "run0")
    os.makedirs(log_dir)
    runner = OnPolicyRunner(env, _train_cfg(), log_dir=log_dir, device="cpu")
    assert runner.alg.storage.observations.shape == (8, 32, 66) and runner.alg.fused is None
    w0 = runner.alg.actor_critic.actor[0].weight.detach().clone()
    runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    assert not torch.equal(w0, runner.alg.actor_critic.actor[0].weight)
    log = runner.last_log
    assert np.isfinite(log["value_loss"]) and np.isfinite(log["surrogate_loss"]) and 1e-5 <= log["learning_rate"] <= 1e-2
    assert log["episode"] == {"rew_x": 0.5} and log["total_timesteps"] == 2 * 8 * 32
    files = sorted(f for f in os.listdir(log_dir) if f.startswith("model_"))
    assert files == ["model_0.pt", "model_1.pt", "model_2.pt"]
    ck = torch.load(os.path.join(log_dir, "model_2.pt"), weights_only=False)
    assert set(ck.keys()) == {"model_state_dict", "optimizer_state_dict", "iter", "infos"} and ck["iter"] == 2
    from nightmare_rl_b200.envs.helpers import get_load_path
    assert get_load_path(str(tmp_path), load_run=-1, checkpoint=-1).endswith("model_2.pt")
    r2 = OnPolicyRunner(FakeEnv(seed=1), _train_cfg(), log_dir=None, device="cpu")
    r2.load(os.path.join(log_dir, "model_2.pt"))
    assert r2.current_learning_iteration == 2
    for a, b in zip(runner.alg.actor_critic.parameters(), r2.alg.actor_critic.parameters()):
        assert torch.equal(a, b)
    pol = r2.get_inference_policy()
    assert pol(torch.zeros(3, 66)).shape == (3, 18)
    # what play.py:65-72 does with the checkpoint
    from rsl_rl.modules import ActorCritic
    nn_ = ActorCritic(66, 66, 18, **_train_cfg()["policy"])
    nn_.load_state_dict(ck["model_state_dict"])


def test_timeout_bootstrap_and_ring():
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    from nightmare_rl_b200.ppo.runner import _EpisodeRing
    ac = ActorCritic(4, 4, 2, actor_hidden_dims=[8], critic_hidden_dims=[8])
    alg = PPO(ac, gamma=0.9, device="cpu", fused_rollout=False)
    alg.init_storage(3, 2, [4], [None], [2])
    obs = torch.randn(3, 4)
    alg.act(obs, obs)
    v = alg.transition.values.clone()
    alg.process_env_step(torch.tensor([1.0, 2.0, 3.0]), torch.tensor([0, 1, 1]), {"time_outs": torch.tensor([0.0, 1.0, 0.0])})
    assert torch.allclose(alg.storage.rewards[0, :, 0], torch.tensor([1.0, 2.0, 3.0]) + 0.9 * v[:, 0] * torch.tensor([0.0, 1.0, 0.0]))
    ring = _EpisodeRing(4, torch.device("cpu"))
    ring.push(torch.tensor([True, False, True]), torch.tensor([1.0, 9.0, 3.0]), torch.tensor([10.0, 99.0, 30.0]))
    ring.push(torch.tensor([False, True, False]), torch.tensor([0.0, 5.0, 0.0]), torch.tensor([0.0, 50.0, 0.0]))
    r, l, n = ring.means()
    assert n == 3 and abs(r - 3.0) < 1e-6 and abs(l - 30.0) < 1e-6
    for _ in range(3):
        ring.push(torch.tensor([True, True, True]), torch.tensor([7.0, 7.0, 7.0]), torch.tensor([1.0, 1.0, 1.0]))
    assert ring.means() == (7.0, 1.0, 4)                            # only the last 4 episodes remain


def _ddp_worker(rank, world, port, shards, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    torch.manual_seed(100 + rank)                                   # different init per rank: the constructor must broadcast rank 0's
    ac = ActorCritic(6, 6, 3, actor_hidden_dims=[8], critic_hidden_dims=[8])
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=1, schedule="adaptive", learning_rate=1e-3, device="cpu", fused_rollout=False)
    _fill(alg, shards[rank])
    alg.update()
    out[rank] = ([p.detach().clone() for p in ac.parameters()], alg.learning_rate)
    dist.destroy_process_group()


def _fill(alg, shard):
    T, N = shard["rewards"].shape[:2]
    alg.init_storage(N, T, [6], [None], [3])
    st = alg.storage
    ac = alg.actor_critic
    with torch.no_grad():
        for t in range(T):
            obs = shard["obs"][t]
            ac.update_distribution(obs)
            st.observations[t] = obs
            st.actions[t] = shard["actions"][t]
            st.mu[t] = ac.action_mean
            st.sigma[t] = ac.action_std
            st.actions_log_prob[t] = ac.get_actions_log_prob(shard["actions"][t]).unsqueeze(1)
            st.values[t] = ac.evaluate(obs)
        st.rewards.copy_(shard["rewards"]); st.dones.copy_(shard["dones"])
        st.step = T
    alg.compute_returns(shard["last_obs"])


def test_gradient_allreduce_equals_single_process_world2():
    """Two gloo ranks with half the envs each must take exactly the optimiser step one process takes on all envs."""
    import torch.multiprocessing as mp
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    g = torch.Generator().manual_seed(3)
    T, N = 5, 8
    full = dict(obs=torch.randn(T, N, 6, generator=g), actions=torch.randn(T, N, 3, generator=g), rewards=torch.randn(T, N, 1, generator=g),
                dones=(torch.rand(T, N, 1, generator=g) < 0.2).to(torch.uint8), last_obs=torch.randn(N, 6, generator=g))
    shards = [{k: (v[:, :4] if v.dim() == 3 else v[:4]) for k, v in full.items()}, {k: (v[:, 4:] if v.dim() == 3 else v[4:]) for k, v in full.items()}]
    torch.manual_seed(100)                                          # same seed as rank 0
    ac = ActorCritic(6, 6, 3, actor_hidden_dims=[8], critic_hidden_dims=[8])
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=1, schedule="adaptive", learning_rate=1e-3, device="cpu", fused_rollout=False)
    _fill(alg, full)
    alg.update()
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_ddp_worker, args=(2, port, shards, out), nprocs=2, join=True)
    for r in (0, 1):
        params, lr = out[r]
        assert lr == alg.learning_rate
        for a, b in zip(ac.parameters(), params):
            assert torch.allclose(a, b, atol=2e-6), "sharded update differs from the single-process update"


@pytest.mark.skipif(not os.path.exists("/root/reference/train.py"), reason="reference tree not present")
def test_reference_train_script_runs_on_the_shims(tmp_path, monkeypatch):
    """The reference's own train.py, byte for byte, against this repo's `rsl_rl` / `envs` import shims.
    No GPU here: the env class is replaced by the CPU stand-in and `learn` is cut to two iterations; what is exercised is
    every import, constructor signature, attribute and call the script makes (train.py:1-54)."""
    import runpy

    import envs.nightmare_v3_env as env_mod
    import rsl_rl.runners as runners_mod
    from nightmare_rl_b200.envs.nightmare_v3_config import NightmareV3Config

    made = {}

    class StandIn(FakeEnv):
        def __init__(self, cfg, log_dir=None, num_threads=1):
            super().__init__(n=cfg.env.num_envs, num_obs=cfg.env.num_obs, num_actions=cfg.env.num_actions)
            made.update(cfg=cfg, log_dir=log_dir, num_threads=num_threads)

    real_learn = runners_mod.OnPolicyRunner.learn

    def short_learn(self, num_learning_iterations, init_at_random_ep_len=False):
        made.update(iters=num_learning_iterations, rand=init_at_random_ep_len)
        return real_learn(self, 2, init_at_random_ep_len)

    monkeypatch.setattr(env_mod, "NightmareV3Env", StandIn)
    monkeypatch.setattr(runners_mod.OnPolicyRunner, "learn", short_learn)
    monkeypatch.setattr(NightmareV3Config, "rl_device", "cpu")
    monkeypatch.setattr(sys, "argv", ["train.py", "-e", "24", "-n", "3"])
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(ROOT)                          # our shims shadow the reference's own `envs` package
    for k in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
        pass                                                    # already ours (imported above)
    runpy.run_path("/root/reference/train.py", run_name="__main__")
    assert made["cfg"].env.num_envs == 24 and made["num_threads"] == 3 and made["rand"] is True and made["iters"] == 1000000000
    runs = os.listdir(tmp_path / "logs" / "nightmare_v3")
    assert len(runs) == 1 and any(f.startswith("model_") for f in os.listdir(tmp_path / "logs" / "nightmare_v3" / runs[0]))
