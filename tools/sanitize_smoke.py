"""Small end-to-end run for compute-sanitizer: env step (both kernel variants), host zero-copy step, reset_idx, DR, policy
kernels (both engines), rollout store."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from envs.nightmare_v3_config import NightmareV3Config
from envs.nightmare_v3_env import NightmareV3Env
from nightmare_rl_b200.ppo import PPO, ActorCritic
dev = torch.device("cuda:0")
for n in ([int(x) for x in sys.argv[1:]] or (301, 4000, 4800, 9472)):      # small / one-wave 7-warp / multi-round 7-warp / 8-warp launch shapes
    cfg = NightmareV3Config(); cfg.env.num_envs = n; cfg.viewer.render = cfg.viewer.record_states = False
    env = NightmareV3Env(cfg, seed=1)
    env.reset()
    env.set_domain_randomization()
    env.episode_length_buf = torch.randint(1200, 1250, (n,))
    for t in range(4):
        env.step(torch.randn(n, 18, device=dev))
    env.step_host(torch.randn(n, 18).pin_memory())
    env.reset_idx([0, 3, n - 1])
    for eng in ("tc5", "mma"):
        os.environ["NM_POLICY_ENGINE"] = eng
        ac = ActorCritic(66, 66, 18, actor_hidden_dims=[54, 42, 30], critic_hidden_dims=[54, 42, 30])
        alg = PPO(ac, device="cuda:0", fused_rollout=True, graph_update=False, seed=1)
        alg.init_storage(n, 3, [66], [None], [18])
        alg.attach_episode_stats(torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(100, device=dev), torch.zeros(100, device=dev),
                                 torch.zeros(1, dtype=torch.int64, device=dev))
        assert alg.prepare_fast_rollout(env, torch.zeros(32, device=dev))
        for t in range(3):
            alg.fast_rollout_step()
    torch.cuda.synchronize()
    print("ok", n, float(env.obs_buf.abs().max()))
