"""Time of the pieces of one PPO update: nm_ppo_grad alone, one captured mini-batch step, compute_returns."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from nightmare_rl_b200.ppo import PPO, ActorCritic
dev = torch.device("cuda:0")
T = 80
for N in ([int(x) for x in sys.argv[1:]] or [4096, 16384]):
    ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
    alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", entropy_coef=0.0015, device="cuda:0", fused_rollout=False)
    alg.init_storage(N, T, [66], [None], [18])
    st = alg.storage
    st.observations.normal_(); st.actions.normal_(); st.mu.normal_(); st.sigma.fill_(1.0); st.values.normal_(); st.returns.normal_()
    st.advantages.normal_(); st.actions_log_prob.fill_(-25.0); st.rewards.normal_(); st.step = T
    fg = alg.fused_grad
    n = T * N // 4
    idx = torch.randperm(T * N, device=dev)[:n].contiguous()
    f = lambda t: t.flatten(0, 1)
    args = (n, idx, f(st.observations), f(st.observations), f(st.actions), f(st.actions_log_prob), f(st.mu), f(st.sigma), f(st.advantages),
            f(st.returns), f(st.values), 0.2, 1.0, 0.0015, True)
    for _ in range(3):
        fg(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        fg(*args)
    e1.record(); torch.cuda.synchronize()
    print(f"N={N}: nm_ppo_grad (n={n}) {e0.elapsed_time(e1) / 20 * 1e3:.0f} us per mini-batch")
    lr_t, acc = torch.tensor(1e-3, device=dev), torch.zeros(2, device=dev)
    for _ in range(3):
        fg.adam(n, lr_t, acc, True, 0.01, 1.0, 0.9, 0.999, 1e-8)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        fg.adam(n, lr_t, acc, True, 0.01, 1.0, 0.9, 0.999, 1e-8)
    e1.record(); torch.cuda.synchronize()
    print(f"N={N}: nm_ppo_adam {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
    t0 = time.perf_counter(); alg.compute_returns(torch.zeros(N, 66, device=dev)); torch.cuda.synchronize(); t1 = time.perf_counter()
    st.step = T
    alg.update(); st.step = T
    torch.cuda.synchronize(); t2 = time.perf_counter()
    alg.update(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"N={N}: compute_returns {1e3 * (t1 - t0):.1f} ms, update (20 mini-batches) {1e3 * (t3 - t2):.1f} ms")
