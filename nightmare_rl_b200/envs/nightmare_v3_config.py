"""Configuration classes with the reference's names, nesting and default values
(``envs/nightmare_v3_config.py:4-146``) so that ``train.py`` / ``play.py`` run unchanged.

Only the values are shared with the reference (they ARE the interface); the GPU env snapshots the
scalars it needs into device constants at construction (see ``nightmare_v3_env.py``)."""
import math

from .base_config import BaseConfig

_LEGS = 6
_STANCE = (0.0, math.pi / 5, 0.0)            # coxa, femur, tibia rest angles (config.py:39-44)


class NightmareV3Config(BaseConfig):
    device = "cpu"
    rl_device = "cuda"

    class env:
        model_path = "models/nightmare_v3/mjmodel.xml"
        num_envs = 8192
        num_obs = 66
        num_privileged_obs = 0
        num_actions = 18
        episode_length_s = 20
        send_timeouts = True
        body_name = "base_link"
        tibia_contact_mode = 1               # 0 ignore, 1 penalise, 2 terminate
        tibia_max_contact_force = 2.0
        body_contact_mode = 1
        body_max_contact_force = 2.0
        termination_contact_force = 160.0

    class viewer:
        render = True
        record_states = True

    class control:
        p_gain = 20
        default_pos = [angle for _ in range(_LEGS) for angle in _STANCE]
        decimation = 2
        action_scale = 0.2

    class noise:
        add_noise = False
        noise_level = 0.1

        class noise_scales:
            lin_vel = 1.0
            ang_vel = 1.0
            gravity = 1.0
            dof_pos = 1.0
            dof_vel = 1.0
            height_measurements = 1.0

    class commands:
        resampling_time = 10

        class ranges:
            max_lin_vel_x = 0.5
            max_lin_vel_y = 0.5
            max_ang_vel = 0.8

    class normalization:
        class obs_scales:
            lin_vel = 2.0
            ang_vel = 0.25
            dof_pos = 1.0
            dof_vel = 0.05
            height_measurements = 5.0

        clip_observations = 100.0
        clip_actions = 1.0

    class rewards:
        class scales:
            # active terms
            termination = -200.0
            tracking_lin_vel = 8.0
            tracking_ang_vel = 6.0
            dof_acc = -2.5e-5
            action_rate = -0.02
            body_contact_forces = -5
            default_position = -0.01
            orientation = -5
            # defined but switched off
            lin_vel_z = 0
            ang_vel_xy = 0
            feet_air_time = 0
            torques = 0
            base_height = 0
            feet_contact_forces = 0
            dof_vel = 0
            stand_still = 0
            collision = 0
            feet_stumble = 0

        tracking_sigma = 0.008
        base_height_target = 0.1
        max_contact_force = 10.0


class NightmareV3ConfigPPO(BaseConfig):
    seed = 1
    runner_class_name = "OnPolicyRunner"

    class policy:
        init_noise_std = 1.0
        actor_hidden_dims = [54, 42, 30]
        critic_hidden_dims = [54, 42, 30]
        activation = "elu"

    class algorithm:
        value_loss_coef = 1.0
        use_clipped_value_loss = True
        clip_param = 0.2
        entropy_coef = 0.0015
        num_learning_epochs = 5
        num_mini_batches = 4
        learning_rate = 1.0e-3
        schedule = "adaptive"
        gamma = 0.99
        lam = 0.95
        desired_kl = 0.01
        max_grad_norm = 1.0

    class runner:
        policy_class_name = "ActorCritic"
        algorithm_class_name = "PPO"
        num_steps_per_env = 80
        max_iterations = 1000000000
        save_interval = 50
        experiment_name = "test"
        run_name = ""
        resume = False
        load_run = -1
        checkpoint = -1
        resume_path = None
