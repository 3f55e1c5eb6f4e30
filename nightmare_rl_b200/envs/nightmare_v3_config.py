"""The reference's configuration surface (``envs/nightmare_v3_config.py:4-146``), table driven.

``train.py`` / ``play.py`` and rsl_rl read these objects by attribute path (``cfg.env.num_envs``,
``train_cfg.algorithm.clip_param``, ...) and through ``class_to_dict``; names, nesting and default values therefore ARE the
interface and are the same as the reference's.  They are written down here as one nested table per config and turned into
the nested attribute classes the callers expect by ``_build`` -- a section is a class whose attributes are its fields, and
``BaseConfig.__init__`` instantiates nested sections so that edits on one instance never leak into another.

The GPU env snapshots the scalars it needs into device constants at construction (``nightmare_v3_env.py``, ``envcfg.py``)."""
import math

from .base_config import BaseConfig


class _S(dict):
    """One config section: ``_S(field=value, sub=_S(...))``."""

    def __init__(self, **fields):
        super().__init__(fields)


def _build(name, table, bases=()):
    """Nested attribute classes from a nested table (sections become classes, everything else a class attribute)."""
    return type(name, bases, {key: _build(key, val) if isinstance(val, _S) else val for key, val in table.items()})


_LEGS = 6
_STANCE = (0.0, math.pi / 5, 0.0)            # coxa, femur, tibia rest angles (reference :39-44)

# reward terms: (name, scale).  The eight non-zero ones are what the default training optimises; the zero ones are defined in the
# env (their functions exist) but switched off; `collision` and `feet_stumble` have a scale and no function in the reference.
_REWARD_SCALES = (
    ("termination", -200.0), ("tracking_lin_vel", 8.0), ("tracking_ang_vel", 6.0), ("dof_acc", -2.5e-5), ("action_rate", -0.02),
    ("body_contact_forces", -5), ("default_position", -0.01), ("orientation", -5),
    ("lin_vel_z", 0), ("ang_vel_xy", 0), ("feet_air_time", 0), ("torques", 0), ("base_height", 0), ("feet_contact_forces", 0),
    ("dof_vel", 0), ("stand_still", 0), ("collision", 0), ("feet_stumble", 0),
)

_unit_scales = lambda *names: _S(**{n: 1.0 for n in names})

NightmareV3Config = _build("NightmareV3Config", _S(
    device="cpu",
    rl_device="cuda",
    env=_S(model_path="models/nightmare_v3/mjmodel.xml", num_envs=8192, num_obs=66, num_privileged_obs=0, num_actions=18,
           episode_length_s=20, send_timeouts=True, body_name="base_link",
           # contact modes: 0 ignore, 1 penalise, 2 terminate
           tibia_contact_mode=1, tibia_max_contact_force=2.0, body_contact_mode=1, body_max_contact_force=2.0,
           termination_contact_force=160.0),
    viewer=_S(render=True, record_states=True),
    control=_S(p_gain=20, default_pos=[angle for _ in range(_LEGS) for angle in _STANCE], decimation=2, action_scale=0.2),
    noise=_S(add_noise=False, noise_level=0.1,
             noise_scales=_unit_scales("lin_vel", "ang_vel", "gravity", "dof_pos", "dof_vel", "height_measurements")),
    commands=_S(resampling_time=10, ranges=_S(max_lin_vel_x=0.5, max_lin_vel_y=0.5, max_ang_vel=0.8)),
    normalization=_S(obs_scales=_S(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, height_measurements=5.0),
                     clip_observations=100.0, clip_actions=1.0),
    rewards=_S(scales=_S(**dict(_REWARD_SCALES)), tracking_sigma=0.008, base_height_target=0.1, max_contact_force=10.0),
), (BaseConfig,))

_HIDDEN = (54, 42, 30)

NightmareV3ConfigPPO = _build("NightmareV3ConfigPPO", _S(
    seed=1,
    runner_class_name="OnPolicyRunner",
    policy=_S(init_noise_std=1.0, actor_hidden_dims=list(_HIDDEN), critic_hidden_dims=list(_HIDDEN), activation="elu"),
    algorithm=_S(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.0015, num_learning_epochs=5,
                 num_mini_batches=4, learning_rate=1.0e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01,
                 max_grad_norm=1.0),
    runner=_S(policy_class_name="ActorCritic", algorithm_class_name="PPO", num_steps_per_env=80, max_iterations=1000000000,
              save_interval=50, experiment_name="test", run_name="", resume=False, load_run=-1, checkpoint=-1, resume_path=None),
), (BaseConfig,))
