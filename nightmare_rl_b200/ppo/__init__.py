"""GPU-resident PPO with the public surface of rsl_rl v1.0.2 (the runner the reference's ``train.py:1,40-54`` and
``play.py:11-12,65-72`` import but does not vendor; pinned by ``setup_vast_ai.sh:24-27``).

``OnPolicyRunner``, ``PPO``, ``ActorCritic`` and ``RolloutStorage`` keep rsl_rl's constructor arguments, method names,
tensor shapes and checkpoint layout (``model_<it>.pt`` = ``{'model_state_dict','optimizer_state_dict','iter','infos'}``),
so the reference's entry scripts and checkpoints work with them.  What is different underneath: the rollout forward pass
is one fused tensor-core kernel (``csrc/nm_policy.cu``), statistics never force a host sync per step, and gradients are
all-reduced over NCCL when ``torch.distributed`` is initialised (environments sharded one process per GPU)."""
from .actor_critic import ActorCritic
from .ppo import PPO
from .runner import OnPolicyRunner
from .storage import RolloutStorage

__all__ = ["ActorCritic", "PPO", "OnPolicyRunner", "RolloutStorage"]
