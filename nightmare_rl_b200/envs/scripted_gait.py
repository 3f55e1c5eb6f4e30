"""Scripted-gait action source: replays joint-target sequences of the reference's nikengine gait engine through the env.

The reference drives the robot with `EngineNode.update(lin, ang, 'awake', 'walk')` in its keyboard player
(custom_play.py:69-74: joint targets, rate-limited to 0.08 rad per step, `ctrl = (target - qpos[-18:]) * kp`).  The gait
engine itself is an interactive tool and is not rebuilt here; its output for a fixed command schedule is a
committed fixture (`tests/golden/nikengine_gait_targets.npz`, written by tools/make_gait_golden.py from the reference's own
code).  This class turns such a sequence into `NightmareV3Env.step` actions -- a deterministic, realistic walking workload
(tripod contacts, stance/swing switching) next to the random-action one.

The env's PD law (reference envs/nightmare_v3_env.py:152-188) is `ctrl = ((clip(a * action_scale, -1, 1) - default_dof_pos)
- dof_pos) * p_gain`, so the joint target theta is reached with `a = (theta + default_dof_pos) / action_scale`.
"""
import numpy as np
import torch


class ScriptedGait:
    def __init__(self, path, num_envs, device, action_scale=0.2, default_dof_pos=None, phase_shift=0):
        """`phase_shift`: env i starts `i * phase_shift` steps into the sequence (wraps inside the walking part), so a
        large batch is not in lockstep."""
        z = np.load(path)
        self.targets = torch.as_tensor(z["targets"], dtype=torch.float32, device=device)              # [T, 18] joint angles
        self.commands = torch.as_tensor(z["commands"], dtype=torch.float32, device=device)
        default = np.array([0.0, np.pi / 5, 0.0] * 6) if default_dof_pos is None else np.asarray(default_dof_pos)
        self.default = torch.as_tensor(default, dtype=torch.float32, device=device)
        self.action_scale = float(action_scale)
        lim = (self.targets + self.default).abs().max().item()
        if lim > 1.0:
            raise ValueError(f"gait targets leave the env's clip range (|theta + default| = {lim:.3f} > 1)")
        self.num_envs = num_envs
        self.offset = (torch.arange(num_envs, device=device) * int(phase_shift))
        self.T = self.targets.shape[0]
        walk = np.flatnonzero(np.abs(z["commands"]).sum(1) > 0)
        self.loop = (int(walk[0]), int(walk[-1]) + 1) if len(walk) else (0, self.T)
        self.t = 0

    def index(self, t):
        """Row of the sequence each env plays at step t: straight through once, then looping over the walking part."""
        i = self.offset + t
        lo, hi = self.loop
        return torch.where(i < self.T, i, lo + (i - self.T) % (hi - lo))

    def actions(self, t=None):
        """[num_envs, 18] actions for env.step at step t (default: internal counter, advanced by one)."""
        if t is None:
            t = self.t
            self.t += 1
        return (self.targets[self.index(t)] + self.default) / self.action_scale

    def joint_targets(self, t):
        return self.targets[self.index(t)]
