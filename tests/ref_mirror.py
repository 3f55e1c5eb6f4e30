"""Numpy transcription of the reference env layer (envs/nightmare_v3_env.py:145-371,399-497) used to
check the oracle's C restatement of it.  Physics (mj_step) is delegated to the oracle's raw physics
entry point; everything else follows the reference statement by statement, in fp64, including its
quirks (stale dof buffers, reset before rewards, alphabetical reward order).  The only substitution is
the RNG: the reference draws from the unseeded global numpy generator (:327-330), this project
defines a Philox stream keyed by (seed, env, step, phase) instead."""
import numpy as np

from nightmare_rl_b200.envs.helpers import class_to_dict
from oracle import oracle as O


class MirrorEnv:
    def __init__(self, cfg, oracle_model, seed=0):
        self.cfg, self.n, self.seed = cfg, cfg.env.num_envs, seed
        self.batch = O.OracleBatch(oracle_model, self.n)
        n = self.n
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        self.dof_pos, self.dof_vel = np.zeros((n, 18)), np.zeros((n, 18))
        self.torques = np.zeros((n, 18))
        self.prev_actions, self.actions = np.zeros((n, 18)), np.zeros((n, 18))
        self.base_heights = np.zeros(n)
        self.feet_contact_forces, self.tibia_contact_forces, self.body_contact_force = np.zeros((n, 6)), np.zeros((n, 6)), np.zeros(n)
        self.reward_scales = class_to_dict(cfg.rewards.scales)
        self.obs_scales = cfg.normalization.obs_scales
        self.rew_buf = np.zeros(n)
        self.reset_buf = np.ones(n, dtype=np.int64)
        self.episode_length_buf = np.zeros(n, dtype=np.int64)
        self.time_out_buf = np.zeros(n, dtype=bool)
        self.feet_air_time = np.zeros((n, 6))
        self.last_contacts = np.zeros((n, 6), dtype=bool)
        self.last_contacts_filt = np.zeros((n, 6), dtype=bool)
        self.commands = np.zeros((n, 3))
        self.commands_scale = np.array([self.obs_scales.lin_vel, self.obs_scales.lin_vel, self.obs_scales.ang_vel])
        self.dt = 0.008 * cfg.control.decimation
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.default_dof_pos = np.array(cfg.control.default_pos, dtype=np.float64)
        self.extras = {}
        self.common_step_counter = 0
        for key in list(self.reward_scales.keys()):
            if self.reward_scales[key] == 0:
                self.reward_scales.pop(key)
            else:
                self.reward_scales[key] *= self.dt
        self.reward_names = [k for k in self.reward_scales if k != "termination"]
        self.episode_sums = {k: np.zeros(n) for k in self.reward_scales}

    # ---- RNG substitution
    def _uniform2(self, env_ids, phase):
        out = np.zeros((len(env_ids), 2))
        s = self.seed
        for k, i in enumerate(env_ids):
            r = O.philox4x32(s & 0xFFFFFFFF, int(i), self.common_step_counter & 0xFFFFFFFF, self.common_step_counter >> 32, phase, s >> 32)
            out[k] = (r[:2] >> 8).astype(np.float64) / 16777216.0
        return out

    def _resample_commands(self, env_ids, phase):
        if len(env_ids) == 0:
            return
        u = self._uniform2(env_ids, phase)
        r = self.cfg.commands.ranges
        self.commands[env_ids, 0] = u[:, 0] * 2 * r.max_lin_vel_x - r.max_lin_vel_x
        self.commands[env_ids, 1] = 0
        self.commands[env_ids, 2] = u[:, 1] * 2 * r.max_ang_vel - r.max_ang_vel
        self.commands[env_ids, :2] *= (np.linalg.norm(self.commands[env_ids, :2], axis=1) > 0.02)[:, None]

    def step(self, actions):
        cfg = self.cfg
        self.prev_actions = self.actions
        all_actions = np.asarray(actions, dtype=np.float32) * cfg.control.action_scale          # float32, as with the policy's tensor (:155)
        self.actions = np.clip(all_actions[:, :18], -cfg.normalization.clip_actions, cfg.normalization.clip_actions).astype(np.float64)
        prev_dof_vel = self.dof_vel.copy()
        velocity_command = ((self.actions - self.default_dof_pos) - self.dof_pos) * cfg.control.p_gain
        self.batch.physics_step(velocity_command, cfg.control.decimation)
        self.episode_length_buf += 1
        self.common_step_counter += 1
        qpos, qvel, _ = self.batch.get_state()
        for i in range(self.n):
            q = qpos[i, 3:7] * np.array([1, -1, -1, -1])
            R = _quat2mat(q)
            cvel = self.batch.get(i, "cvel").reshape(-1, 6)[1]
            self.base_lin_vel[i] = R @ cvel[3:]
            self.base_ang_vel[i] = R @ cvel[:3]
            self.projected_gravity[i] = R @ np.array([0, 0, -9.81])
            sd = self.batch.get(i, "sensordata")
            self.tibia_contact_forces[i], self.feet_contact_forces[i], self.body_contact_force[i] = sd[:6], sd[6:12], sd[12]
            self.base_heights[i] = self.batch.get(i, "xipos").reshape(-1, 3)[1, 2]
        self.dof_pos, self.dof_vel = qpos[:, -18:].copy(), qvel[:, -18:].copy()
        self.dof_acc = (self.dof_vel - prev_dof_vel) / self.dt
        self.tibia_contact_forces *= (self.feet_contact_forces == 0)
        ids = np.nonzero(self.episode_length_buf % int(cfg.commands.resampling_time / self.dt) == 0)[0]
        self._resample_commands(ids, 0)
        self.reset_buf = np.zeros_like(self.reset_buf)
        self.time_out_buf = self.episode_length_buf > self.max_episode_length
        self.reset_buf |= self.time_out_buf
        self.reset_buf |= self.feet_contact_forces.max(axis=1) > cfg.env.termination_contact_force
        if cfg.env.tibia_contact_mode == 2:
            self.reset_buf |= self.tibia_contact_forces.max(axis=1) > cfg.env.tibia_max_contact_force
        if cfg.env.body_contact_mode == 2:
            self.reset_buf |= self.body_contact_force > cfg.env.body_max_contact_force
        down = np.array([0, 0, -1])
        with np.errstate(invalid="ignore"):
            ang = np.arccos(self.projected_gravity @ down / np.linalg.norm(self.projected_gravity, axis=1))
        self.reset_buf |= ang > 60 * np.pi / 180
        self.reset_idx(np.nonzero(self.reset_buf)[0])
        self.rew_buf[:] = 0.0
        for name in self.reward_names:
            rew = getattr(self, "_reward_" + name)() * self.reward_scales[name]
            self.rew_buf += rew
            self.episode_sums[name] += rew
        if "termination" in self.reward_scales:
            rew = (self.reset_buf * ~self.time_out_buf) * self.reward_scales["termination"]
            self.rew_buf += rew
            self.episode_sums["termination"] += rew
        s = self.obs_scales
        obs = np.concatenate((self.base_lin_vel * s.lin_vel, self.base_ang_vel * s.ang_vel, self.projected_gravity,
                              self.commands[:, :3] * self.commands_scale, (self.dof_pos - self.default_dof_pos) * s.dof_pos,
                              self.dof_vel * s.dof_vel, self.actions), axis=-1)
        obs = np.clip(obs, -cfg.normalization.clip_observations, cfg.normalization.clip_observations)
        return obs.astype(np.float32), self.rew_buf.astype(np.float32), self.reset_buf.copy()

    def reset_idx(self, env_ids):
        if len(env_ids) == 0:
            return
        qpos, qvel, _ = self.batch.get_state()
        qpos[env_ids] = np.r_[0, 0, 0.15, 1, 0, 0, 0, np.zeros(18)]
        qvel[env_ids] = 0
        self.batch.set_state(qpos, qvel, None)
        self._resample_commands(env_ids, 1)
        self.feet_air_time[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
        self.extras["episode"] = {}
        for key in self.episode_sums:
            self.extras["episode"]["rew_" + key] = np.mean(self.episode_sums[key][env_ids]) / self.max_episode_length_s
            self.episode_sums[key][env_ids] = 0.0
        self.extras["time_outs"] = self.time_out_buf.astype(np.float32)

    # ---- reward terms (:399-497)
    def _reward_lin_vel_z(self): return np.square(self.base_lin_vel[:, 2])
    def _reward_ang_vel_xy(self): return np.sum(np.square(self.base_ang_vel[:, :2]), axis=1)
    def _reward_orientation(self): return np.sum(np.square(self.projected_gravity[:, :2]), axis=1)
    def _reward_base_height(self): return np.square(self.base_heights - self.cfg.rewards.base_height_target)
    def _reward_torques(self): return np.sum(np.square(self.torques), axis=1)
    def _reward_dof_vel(self): return np.sum(np.square(self.dof_vel), axis=1)
    def _reward_dof_acc(self): return np.sum(np.square(self.dof_acc), axis=1)
    def _reward_action_rate(self): return np.sum(np.square(self.prev_actions - self.actions), axis=1)
    def _reward_tracking_lin_vel(self):
        return np.exp(-np.sum(np.square(self.commands[:, :2] - self.base_lin_vel[:, :2]), axis=1) / self.cfg.rewards.tracking_sigma)
    def _reward_tracking_ang_vel(self):
        return np.exp(-np.square(self.commands[:, 2] - self.base_ang_vel[:, 2]) / self.cfg.rewards.tracking_sigma)
    def _reward_feet_air_time(self):
        contact = self.feet_contact_forces > 1.0
        filt = np.logical_or(contact, self.last_contacts)
        self.feet_air_time += self.dt
        self.feet_air_time *= filt == self.last_contacts_filt
        self.last_contacts, self.last_contacts_filt = contact, filt
        t = self.feet_air_time
        single = (t > 1.0) * (t - 1.0) + (t < 0.5) * (0.5 - t)
        return np.sum(np.square(single), axis=1)
    def _reward_body_contact_forces(self):
        rew = np.zeros(self.n)
        if self.cfg.env.tibia_contact_mode == 1: rew += np.sum(self.tibia_contact_forces, axis=1)
        if self.cfg.env.body_contact_mode == 1: rew += self.body_contact_force
        return rew
    def _reward_stand_still(self):
        return np.sum(np.abs(self.dof_pos - self.default_dof_pos), axis=1) * (np.linalg.norm(self.commands[:, :2], axis=1) < 0.01)
    def _reward_feet_contact_forces(self):
        m = self.cfg.rewards.max_contact_force
        return np.sum(((self.feet_contact_forces - m) * (self.feet_contact_forces > m)) ** 2, axis=1)
    def _reward_default_position(self): return np.sum(np.square(self.dof_pos - self.default_dof_pos), axis=1)


def _quat2mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])
