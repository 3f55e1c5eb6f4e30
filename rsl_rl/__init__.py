"""Import shim: the reference's entry scripts do ``from rsl_rl.runners import OnPolicyRunner`` (train.py:1) and
``from rsl_rl.modules import ActorCritic`` (play.py:12); rsl_rl v1.0.2 itself is not vendored by the reference and is not
installable offline.  The implementation lives in ``nightmare_rl_b200.ppo``."""
