"""Kernel-time breakdown of one PPO update (eager path) with torch.profiler."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from nightmare_rl_b200.ppo import PPO, ActorCritic
dev = torch.device("cuda:0")
T, N = 80, 4096
ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", entropy_coef=0.0015, device="cuda:0", fused_rollout=False, graph_update=False)
alg.init_storage(N, T, [66], [None], [18])
st = alg.storage
st.observations.normal_(); st.actions.normal_(); st.mu.normal_(); st.sigma.fill_(1.0); st.values.normal_(); st.returns.normal_(); st.advantages.normal_(); st.actions_log_prob.fill_(-25.0); st.step = T
alg.update(); st.step = T
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    alg.update()
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev if e.device_time_total and e.device_type.name == "CUDA") or sum(e.self_device_time_total for e in ev)
rows = sorted(ev, key=lambda e: -e.self_device_time_total)[:22]
print(f"total self device time {sum(e.self_device_time_total for e in ev) / 1e3:.1f} ms for 20 mini-batches")
for e in rows:
    print(f"{e.self_device_time_total / 1e3:8.2f} ms  n={e.count:5d}  {e.key[:90]}")
