"""Fused tensor-core policy forward (nm_policy_act ≙ rsl_rl PPO.act: actor(obs), critic(obs), Normal sample + log-prob;
reference call sites train.py:54 / play.py:122) against the plain PyTorch fp32 modules, and the PPO runner on the real
CUDA environment."""
import os

import numpy as np
import pytest
import torch

from conftest import NMB

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _ac(seed=0):
    from nightmare_rl_b200.ppo import ActorCritic
    torch.manual_seed(seed)
    ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30], activation="elu", init_noise_std=1.0).to(DEV)
    with torch.no_grad():
        ac.std.copy_(torch.linspace(0.3, 1.5, 18))
    return ac


@pytest.mark.parametrize("engine", ["tc5", "mma"])
@pytest.mark.parametrize("n", [1, 15, 64, 1000, 4096])
def test_policy_kernel_matches_torch_fp32(n, engine):
    from nightmare_rl_b200.ppo.policy_kernel import FusedPolicy
    ac = _ac()
    fp = FusedPolicy(ac, DEV, seed=5, engine=engine)
    assert fp.engine == engine
    obs = torch.randn(n, 66, device=DEV) * 2.0
    actions, mean, value, logp = fp.act(obs, step=3)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref_mean, ref_value = ac.actor(obs), ac.critic(obs)[:, 0]
    # 3xTF32 tensor-core MMAs with fp32 accumulation vs the fp32 torch modules: 1e-4 absolute on O(1) outputs
    assert (mean - ref_mean).abs().max() < 1e-4 and (value - ref_value).abs().max() < 1e-4
    assert actions.shape == (n, 18) and logp.shape == (n,) and torch.isfinite(actions).all()
    # log-prob is consistent with the kernel's own mean/std and the sampled action (fp32 arithmetic)
    lp = torch.distributions.Normal(mean, ac.std.detach().expand_as(mean)).log_prob(actions).sum(-1)
    assert (lp - logp).abs().max() < 2e-4
    # reproducible in (seed, step); different step -> different noise; deterministic -> mean
    a2, _, _, _ = fp.act(obs, step=3)
    a3, _, _, _ = fp.act(obs, step=4)
    ad, md, _, _ = fp.act(obs, step=3, deterministic=True)
    assert torch.equal(actions, a2) and not torch.equal(actions, a3) and torch.equal(ad, md)


@pytest.mark.parametrize("engine", ["tc5", "mma"])
def test_policy_noise_is_standard_normal_and_weights_reload(engine):
    from nightmare_rl_b200.ppo.policy_kernel import FusedPolicy
    ac = _ac(1)
    fp = FusedPolicy(ac, DEV, seed=9, engine=engine)
    obs = torch.randn(8192, 66, device=DEV)
    actions, mean, _, _ = fp.act(obs, step=1)
    z = ((actions - mean) / ac.std.detach()).flatten()
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z ** 3).mean())) < 0.05 and abs(float((z ** 4).mean()) - 3.0) < 0.1
    zz = z.view(8192, 18)
    assert abs(float(torch.corrcoef(zz.T)[0, 1])) < 0.05            # independent across action columns
    with torch.no_grad():
        for p in ac.parameters():
            p.mul_(0.5)
    fp.load(ac)
    _, mean2, value2, _ = fp.act(obs, step=1)
    with torch.no_grad():
        assert (mean2 - ac.actor(obs)).abs().max() < 1e-4 and (value2 - ac.critic(obs)[:, 0]).abs().max() < 1e-4


def test_runner_on_cuda_env(tmp_path):
    """train.py's flow (reference train.py:29-54) for 2 iterations on 256 GPU envs: finite losses, checkpoint written,
    fused rollout kernel actually used."""
    from envs.helpers import class_to_dict
    from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
    from envs.nightmare_v3_env import NightmareV3Env
    from rsl_rl.runners import OnPolicyRunner
    cfg, tc = NightmareV3Config(), NightmareV3ConfigPPO()
    cfg.env.num_envs = 256
    cfg.env.model_path = NMB
    cfg.viewer.render = cfg.viewer.record_states = False
    tc.runner.num_steps_per_env = 16
    log_dir = str(tmp_path / "run")
    os.makedirs(log_dir)
    env = NightmareV3Env(cfg, log_dir=log_dir, num_threads=1)
    runner = OnPolicyRunner(env, class_to_dict(tc), log_dir=log_dir, device=cfg.rl_device)
    runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    log = runner.last_log
    assert np.isfinite(log["value_loss"]) and np.isfinite(log["surrogate_loss"])
    assert runner.alg.fused is not None and runner.alg.fused.launches >= 2 * 16
    assert set(log["episode"].keys()) >= {"rew_tracking_lin_vel", "rew_termination"}
    assert os.path.exists(os.path.join(log_dir, "model_2.pt"))
    # first PPO epoch: the fused rollout log-probs must agree with the autograd path's (ratio ~ 1)
    obs = env.get_observations()
    ac = runner.alg.actor_critic
    runner.alg.fused.load(ac)                                       # what PPO.act does lazily after an update
    a, m, v, lp = runner.alg.fused.act(obs, step=12345)
    with torch.no_grad():
        ac.update_distribution(obs)
        lp_ref = ac.get_actions_log_prob(a)
    assert (lp - lp_ref).abs().max() < 2e-3


def test_graph_replayed_update_equals_eager_update():
    """PPO.update as a replayed CUDA graph (device-side KL-adaptive learning rate) == the eager rsl_rl update rule."""
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    T, N = 16, 512
    g = torch.Generator(device=DEV).manual_seed(0)
    data = dict(obs=torch.randn(T, N, 66, device=DEV, generator=g), actions=torch.randn(T, N, 18, device=DEV, generator=g),
                rewards=torch.randn(T, N, 1, device=DEV, generator=g) * 0.1, dones=(torch.rand(T, N, 1, device=DEV, generator=g) < 0.05).to(torch.uint8),
                last=torch.randn(N, 66, device=DEV, generator=g))
    results = []
    for graphed in (False, True):
        torch.manual_seed(11)
        ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
        alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0, entropy_coef=0.0015,
                  learning_rate=1e-3, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01, device="cuda:0", fused_rollout=False, graph_update=graphed)
        alg.init_storage(N, T, [66], [None], [18])
        for it in range(3):
            st = alg.storage
            with torch.no_grad():
                for t in range(T):
                    ac.update_distribution(data["obs"][t])
                    st.observations[t] = data["obs"][t]; st.actions[t] = data["actions"][t]
                    st.mu[t] = ac.action_mean; st.sigma[t] = ac.action_std
                    st.actions_log_prob[t] = ac.get_actions_log_prob(data["actions"][t]).unsqueeze(1)
                    st.values[t] = ac.evaluate(data["obs"][t])
                st.rewards.copy_(data["rewards"]); st.dones.copy_(data["dones"]); st.step = T
            alg.compute_returns(data["last"])
            torch.manual_seed(100 + it)                                  # same mini-batch permutation in both modes
            vl, sl = alg.update()
        results.append(([p.detach().clone() for p in ac.parameters()], alg.learning_rate, vl, sl))
    (pe, lre, vle, sle), (pg, lrg, vlg, slg) = results
    assert abs(lre - lrg) < 1e-9 * max(1.0, lre) + 1e-9 and lre != 1e-3          # the schedule moved, identically
    assert abs(vle - vlg) < 1e-3 * max(1.0, abs(vle)) and abs(sle - slg) < 1e-4
    for a, b in zip(pe, pg):
        assert torch.allclose(a, b, atol=2e-3, rtol=1e-2)           # 60 Adam steps at lr up to 1e-2 amplify fp32 reassociation


def test_fused_transition_store_equals_torch_path():
    """nm_rollout_store (one launch) == PPO.process_env_step + RolloutStorage.add_transitions + the runner's statistics."""
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    T, N = 6, 700
    algs, stats = [], []
    for fused_store in (True, False):
        torch.manual_seed(3)
        ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
        alg = PPO(ac, gamma=0.99, device="cuda:0", fused_rollout=True, graph_update=False, seed=5)
        alg.init_storage(N, T, [66], [None], [18])
        st = (torch.zeros(N, device=DEV), torch.zeros(N, device=DEV), torch.zeros(100, device=DEV), torch.zeros(100, device=DEV),
              torch.zeros(1, dtype=torch.int64, device=DEV))
        alg.attach_episode_stats(*st)
        if not fused_store:
            alg._store_fused = lambda *a, **k: False          # force the PyTorch path
        algs.append(alg); stats.append(st)
    g = torch.Generator(device=DEV).manual_seed(9)
    for t in range(T):
        obs = torch.randn(N, 66, device=DEV, generator=g)
        rew = torch.randn(N, device=DEV, generator=g)
        done = (torch.rand(N, device=DEV, generator=g) < 0.1).to(torch.int64)
        tout = ((torch.rand(N, device=DEV, generator=g) < 0.5) & (done > 0)).float()
        for alg in algs:
            a = alg.act(obs, obs)
            alg.process_env_step(rew, done, {"time_outs": tout})
        assert torch.equal(algs[0].storage.actions[t], algs[1].storage.actions[t])
    sa, sb = algs[0].storage, algs[1].storage
    for name in ("observations", "actions", "mu", "sigma", "values", "actions_log_prob", "rewards", "dones"):
        assert torch.equal(getattr(sa, name), getattr(sb, name)), name
    assert sa.step == sb.step == T
    (cr0, cl0, rr0, rl0, rc0), (cr1, cl1, rr1, rl1, rc1) = stats
    assert torch.equal(cr0, cr1) and torch.equal(cl0, cl1) and int(rc0) == int(rc1) > 100
    n = min(int(rc0), 100)                                     # ring order differs (atomics); its content within a step does not
    assert abs(float(rl0[:n].sum()) - float(rl1[:n].sum())) <= 60 and torch.isfinite(rr0).all()


def test_prebound_rollout_equals_standard_rollout():
    """The runner's pre-bound rollout (nm_policy_act_store -> nm_step -> nm_rollout_store on fixed pointers) fills the
    rollout buffer exactly like the rsl_rl-style loop act -> env.step -> process_env_step."""
    from envs.nightmare_v3_config import NightmareV3Config
    from envs.nightmare_v3_env import NightmareV3Env
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    T, N = 12, 300
    out = []
    for fast in (False, True):
        cfg = NightmareV3Config()
        cfg.env.num_envs = N
        cfg.env.model_path = NMB
        cfg.viewer.render = cfg.viewer.record_states = False
        env = NightmareV3Env(cfg, seed=3)
        env.reset()
        env.episode_length_buf = torch.arange(N, dtype=torch.int64) % 13 + 1240      # time-outs inside the window
        torch.manual_seed(1)
        ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
        alg = PPO(ac, gamma=0.99, device="cuda:0", fused_rollout=True, graph_update=False, seed=2)
        alg.init_storage(N, T, [66], [None], [18])
        st = (torch.zeros(N, device=DEV), torch.zeros(N, device=DEV), torch.zeros(100, device=DEV), torch.zeros(100, device=DEV),
              torch.zeros(1, dtype=torch.int64, device=DEV))
        alg.attach_episode_stats(*st)
        if fast:
            assert alg.prepare_fast_rollout(env, torch.zeros(32, device=DEV))
            for t in range(T):
                alg.fast_rollout_step()
        else:
            obs = env.get_observations()
            for t in range(T):
                a = alg.act(obs, obs)
                obs, _, rew, done, infos = env.step(a)
                alg.process_env_step(rew, done, infos)
        torch.cuda.synchronize()
        out.append((alg.storage, st, env.get_state()[0].clone()))
    (sa, sta, qa), (sb, stb, qb) = out
    for name in ("observations", "actions", "mu", "sigma", "values", "actions_log_prob", "rewards", "dones"):
        assert torch.equal(getattr(sa, name), getattr(sb, name)), name
    assert torch.equal(qa, qb) and torch.equal(sta[0], stb[0]) and int(sta[4]) == int(stb[4])
    assert sa.dones.sum() > 0


def test_split_k_linear_backward_matches_plain_linear():
    """The split-K weight-gradient GEMM of the update's Linear layers == autograd's plain nn.Linear (fp32)."""
    from nightmare_rl_b200.ppo.actor_critic import _Linear
    torch.manual_seed(0)
    lin = _Linear(66, 54).to(DEV)
    ref = torch.nn.Linear(66, 54).to(DEV)
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(81920, 66, device=DEV, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    w = torch.randn(81920, 54, device=DEV)
    (lin(x) * w).sum().backward()
    (ref(x2) * w).sum().backward()
    assert torch.allclose(lin.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-2)
    assert torch.allclose(lin.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-2)
    assert torch.allclose(x.grad, x2.grad, rtol=1e-4, atol=1e-4)


def test_fused_ppo_head_matches_autograd():
    """nm_ppo_head (loss sums + gradients in one launch) == the eager rsl_rl loss through autograd, incl. clip edges."""
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    torch.manual_seed(0)
    n = 5000
    ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
    alg = PPO(ac, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0015, device="cuda:0", fused_rollout=False, graph_update=False)
    with torch.no_grad():
        ac.std.copy_(torch.linspace(0.5, 1.5, 18))
    g = torch.Generator(device=DEV).manual_seed(1)
    obs = torch.randn(n, 66, device=DEV, generator=g)
    with torch.no_grad():
        ac.update_distribution(obs)
        mu0 = ac.action_mean.clone()
        act = mu0 + ac.std * torch.randn(n, 18, device=DEV, generator=g)
        old_mu = mu0 + 0.05 * torch.randn(n, 18, device=DEV, generator=g)
        old_sigma = (ac.std * (1 + 0.05 * torch.randn(18, device=DEV, generator=g))).expand(n, 18).contiguous()
        old_logp = (ac.get_actions_log_prob(act) + 0.3 * torch.randn(n, device=DEV, generator=g)).unsqueeze(1)      # ratios on both sides of the clip
        v0 = ac.evaluate(obs)
        tgt = v0 + 0.3 * torch.randn(n, 1, device=DEV, generator=g)
        ret = v0 + torch.randn(n, 1, device=DEV, generator=g)
        adv = torch.randn(n, 1, device=DEV, generator=g)
    res = []
    for fused in (False, True):
        alg.fused_head = fused
        for p in ac.parameters():
            p.grad = None
        loss, vl, sl, kl = alg._minibatch_loss(obs, obs, act, tgt, adv, ret, old_logp, old_mu, old_sigma)
        loss.backward()
        res.append((float(loss), float(vl), float(sl), float(kl), [p.grad.clone() for p in ac.parameters()]))
    (l0, v0_, s0, k0, g0), (l1, v1, s1, k1, g1) = res
    assert abs(l0 - l1) < 1e-4 * max(1, abs(l0)) and abs(v0_ - v1) < 1e-4 * max(1, abs(v0_)) and abs(s0 - s1) < 1e-5 and abs(k0 - k1) < 1e-5
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-6), (a - b).abs().max()


@pytest.mark.parametrize("hidden", [[54, 42, 30], [64, 40, 32]])
def test_fused_ppo_grad_matches_autograd(hidden):
    """nm_ppo_grad (gather + both MLPs forward/backward + loss head in one launch) == autograd on the eager rsl_rl loss:
    loss sums to fp32 accuracy (3xTF32 forward), gradients to TF32-backward accuracy; ragged batch (n % 128 != 0)."""
    from nightmare_rl_b200.ppo import PPO, ActorCritic
    torch.manual_seed(0)
    rows, n = 3000, 2477
    # [64, 40, 32]: widths that are multiples of 8 (no zero padding inside the tiles)
    ac = ActorCritic(66, 66, 18, actor_hidden_dims=hidden, critic_hidden_dims=hidden)
    alg = PPO(ac, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0015, device="cuda:0", fused_rollout=False, graph_update=True)
    assert alg.fused_grad is not None
    fg = alg.fused_grad
    with torch.no_grad():
        ac.std.copy_(torch.linspace(0.5, 1.5, 18))
    assert ac.std.data_ptr() == fg.flat.data_ptr()                     # parameters are views of the flat buffer
    g = torch.Generator(device=DEV).manual_seed(1)
    obs = torch.randn(rows, 66, device=DEV, generator=g)
    with torch.no_grad():
        ac.update_distribution(obs)
        mu0 = ac.action_mean.clone()
        act = mu0 + ac.std * torch.randn(rows, 18, device=DEV, generator=g)
        old_mu = mu0 + 0.05 * torch.randn(rows, 18, device=DEV, generator=g)
        old_sigma = (ac.std * (1 + 0.05 * torch.randn(18, device=DEV, generator=g))).expand(rows, 18).contiguous()
        old_logp = (ac.get_actions_log_prob(act) + 0.3 * torch.randn(rows, device=DEV, generator=g)).unsqueeze(1).contiguous()
        v0 = ac.evaluate(obs)
        tgt = (v0 + 0.3 * torch.randn(rows, 1, device=DEV, generator=g)).contiguous()
        ret = (v0 + torch.randn(rows, 1, device=DEV, generator=g)).contiguous()
        adv = torch.randn(rows, 1, device=DEV, generator=g)
    idx = torch.randperm(rows, device=DEV, generator=g)[:n].contiguous()
    out = fg(n, idx, obs, obs, act, old_logp, old_mu, old_sigma, adv, ret, tgt, 0.2, 1.0, 0.0015, True)
    torch.cuda.synchronize()
    sums = (out / n).tolist()
    got = fg.flat_grad.clone()
    # autograd reference on the same rows, fp32 backward
    alg.fused_head = False
    alg.tf32_backward = False
    params = list(ac.parameters())
    loss, vl, sl, kl = alg._minibatch_loss(obs[idx], obs[idx], act[idx], tgt[idx], adv[idx], ret[idx], old_logp[idx], old_mu[idx], old_sigma[idx])
    grads = torch.autograd.grad(loss, params)
    assert abs(sums[0] - float(sl)) < 2e-5 and abs(sums[1] - float(vl)) < 1e-4 * max(1.0, float(vl)) and abs(sums[2] - float(kl)) < 2e-5, (sums, float(sl), float(vl), float(kl))
    worst = 0.0
    for p, gr in zip(params, grads):
        off = (p.data_ptr() - fg.flat.data_ptr()) // 4
        mine = got[off:off + p.numel()].view_as(p)
        scale = gr.abs().max().item()
        err = (mine - gr).abs().max().item() / max(scale, 1e-12)
        worst = max(worst, err)
        assert err < 5e-3, (tuple(p.shape), err, scale)
    print(f"\n[ppo-grad] worst gradient error relative to each tensor's max |grad|: {worst:.2e}")
    # identity rows (idx = None) and an exact multiple of the CTA batch
    out2 = fg(256, None, obs, obs, act, old_logp, old_mu, old_sigma, adv, ret, tgt, 0.2, 1.0, 0.0015, True)
    torch.cuda.synchronize()
    sl2 = alg._minibatch_loss(obs[:256], obs[:256], act[:256], tgt[:256], adv[:256], ret[:256], old_logp[:256], old_mu[:256], old_sigma[:256])[2]
    assert abs(out2[0].item() / 256 - float(sl2)) < 2e-5


def test_fused_gae_matches_loop():
    """nm_gae == the rsl_rl compute_returns loop (returns exactly up to fp32 rounding, normalised advantages to 1e-5)."""
    from nightmare_rl_b200.ppo.storage import RolloutStorage
    T, N = 24, 1000
    g = torch.Generator(device=DEV).manual_seed(3)
    sts = []
    for fused in (False, True):
        st = RolloutStorage(N, T, [66], [None], [18], device=DEV)
        st.fused_gae = fused
        gg = torch.Generator(device=DEV).manual_seed(3)
        st.rewards.copy_(torch.randn(T, N, 1, device=DEV, generator=gg))
        st.values.copy_(torch.randn(T, N, 1, device=DEV, generator=gg))
        st.dones.copy_((torch.rand(T, N, 1, device=DEV, generator=gg) < 0.05).to(torch.uint8))
        last = torch.randn(N, 1, device=DEV, generator=gg)
        st.compute_returns(last, 0.99, 0.95)
        sts.append(st)
    torch.cuda.synchronize()
    assert torch.allclose(sts[0].returns, sts[1].returns, rtol=1e-5, atol=1e-5)
    assert torch.allclose(sts[0].advantages, sts[1].advantages, rtol=1e-4, atol=1e-5)
    assert abs(sts[1].advantages.mean().item()) < 1e-5 and abs(sts[1].advantages.std().item() - 1) < 1e-4


def test_checkpoint_resume_with_fused_update(tmp_path):
    """train.py -r flow (reference train.py:49-52): save, rebuild, load, continue.  With the fused update the parameters and
    their gradients are views of two flat buffers; loading a checkpoint must keep them so (in-place copies) and Adam's state
    must come back with it."""
    from envs.helpers import class_to_dict
    from envs.nightmare_v3_config import NightmareV3Config, NightmareV3ConfigPPO
    from envs.nightmare_v3_env import NightmareV3Env
    from rsl_rl.runners import OnPolicyRunner

    def make(sub):
        cfg, tc = NightmareV3Config(), NightmareV3ConfigPPO()
        cfg.env.num_envs = 256
        cfg.env.model_path = NMB
        cfg.viewer.render = cfg.viewer.record_states = False
        tc.runner.num_steps_per_env = 8
        log_dir = str(tmp_path / sub)
        os.makedirs(log_dir)
        env = NightmareV3Env(cfg, log_dir=log_dir, num_threads=1)
        return OnPolicyRunner(env, class_to_dict(tc), log_dir=log_dir, device=cfg.rl_device)

    r1 = make("a")
    r1.learn(num_learning_iterations=3, init_at_random_ep_len=True)
    assert r1.alg.fused_grad is not None
    path = str(tmp_path / "ck.pt")
    r1.save(path)
    saved = {k: v.clone() for k, v in r1.alg.actor_critic.state_dict().items()}
    exp_avg = r1.alg.optimizer.state[r1.alg.actor_critic.std]["exp_avg"].clone()
    lr1 = r1.alg.learning_rate
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck.keys()) == {"model_state_dict", "optimizer_state_dict", "iter", "infos"}
    assert isinstance(ck["optimizer_state_dict"]["param_groups"][0]["lr"], float)
    assert set(ck["model_state_dict"].keys()) == {"std"} | {f"{net}.{i}.{w}" for net in ("actor", "critic") for i in (0, 2, 4, 6) for w in ("weight", "bias")}
    r2 = make("b")
    r2.load(path)
    fg = r2.alg.fused_grad
    lo, hi = fg.flat.data_ptr(), fg.flat.data_ptr() + 4 * fg.flat.numel()
    for k, p in r2.alg.actor_critic.named_parameters():
        assert lo <= p.data_ptr() < hi and p.grad is not None and fg.flat_grad.data_ptr() <= p.grad.data_ptr() < fg.flat_grad.data_ptr() + 4 * fg.flat_grad.numel()
        assert torch.equal(p.detach(), saved[k])
    assert torch.equal(r2.alg.optimizer.state[r2.alg.actor_critic.std]["exp_avg"], exp_avg)
    assert abs(r2.alg.learning_rate - lr1) < 1e-12 and r2.current_learning_iteration == 3
    std_state = r2.alg.optimizer.state[r2.alg.actor_critic.std]
    step1 = float(std_state["step"])
    n_upd = r2.alg.num_learning_epochs * r2.alg.num_mini_batches
    assert step1 == 3 * n_upd
    r2.learn(num_learning_iterations=1)
    # Adam's state CONTINUES from the checkpoint through the graph capture's warm-up steps: the step counter advanced by one
    # iteration's updates (not reset to them), and the first moment is the old one decayed, not a fresh one
    std_state = r2.alg.optimizer.state[r2.alg.actor_critic.std]
    assert float(std_state["step"]) == step1 + n_upd
    sq = r2.alg.optimizer.state[r2.alg.actor_critic.std]["exp_avg_sq"]
    assert (sq > 0).all()
    r2.learn(num_learning_iterations=1)
    assert np.isfinite(r2.last_log["value_loss"]) and r2.current_learning_iteration == 5
    moved = max((p.detach() - saved[k]).abs().max().item() for k, p in r2.alg.actor_critic.named_parameters())
    assert 0 < moved < 0.5
