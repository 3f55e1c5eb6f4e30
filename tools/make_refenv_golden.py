#!/usr/bin/env python
"""Golden vectors from the reference's OWN environment code, run over this repository's physics.

`envs/nightmare_v3_env.py` of the reference cannot be imported as is: its first line imports `mujoco`, which is not
installable here.  Everything ELSE in that file -- the action map, the PD law, the buffer updates and their
ordering, command resampling, termination, reset, the reward terms and their order, the observation vector, the extras
(rows E1-E18 and quirks Q1-Q12 of SURVEY.md §8a) -- is plain numpy/torch.  This script therefore injects a minimal stand-in
for the seven MuJoCo entry points the file uses (`MjModel.from_xml_path`, `MjData`, `mj_name2id`, `mj_step`, `mju_negQuat`,
`mju_rotVecQuat`, `mjtObj`; reference envs/nightmare_v3_env.py:37,38,48,200,216-219) whose `mj_step` is the CPU oracle's
physics (oracle/nm_oracle.c), imports the UNMODIFIED reference module from the read-only tree, and records what the
reference's `NightmareV3Env.reset()/step()` return for seeded action sequences.

What the resulting fixture pins: the ENV LAYER of the oracle (and, through the oracle, of the CUDA kernel) against the
reference's actual code, given identical physics.  What it does not pin: the physics itself (still unpinned against MuJoCo).

The one substitution besides physics is the random source.  The reference draws command samples from the unseeded global
`np.random.rand` (:327-330); this project defines them as Philox4x32-10 uniforms keyed by (seed, env id, step counter,
phase).  While the reference code runs, `np.random.rand` is replaced by a function that returns exactly those uniforms for
the env ids being resampled, so both sides see the same commands.

    python tools/make_refenv_golden.py [/root/reference]      -> tests/golden/reference_env_on_oracle_physics.npz
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

NMB = os.path.join(ROOT, "models", "nightmare_v3", "mjmodel.nmb")
SEED = 7


# ------------------------------------------------------------------------------------------------ the MuJoCo stand-in
def make_mujoco_stub(nmb_path):
    from nightmare_rl_b200 import mjcf
    cm = mjcf.CompiledModel.load(nmb_path)
    om = O.OracleModel(nmb_path)
    mj = types.ModuleType("mujoco")

    class _Opt:
        timestep = float(cm.arrays["opt_real"][0])

    class MjModel:
        def __init__(self):
            self.nv, self.nq, self.ngeom, self.nbody = cm.nv, cm.nq, cm.ngeom, cm.nbody
            self.opt = _Opt()
            self.qpos0 = np.array(cm.qpos0, dtype=np.float64)

        @classmethod
        def from_xml_path(cls, path):                           # the .nmb next to the MJCF is its compiled form
            return cls()

    class MjData:
        """State lives in numpy arrays, like the binding's views; assignments copy INTO them (`data.qvel = 0` broadcasts)."""

        def __init__(self, model):
            self._b = O.OracleBatch(om, 1)
            q, v, _ = self._b.get_state()
            self._qpos, self._qvel = q[0].copy(), v[0].copy()
            self._ctrl = np.zeros(cm.nu)
            self.qfrc_applied = np.zeros(cm.nv)                 # nothing in the reference ever writes it (quirk Q7)
            self.act = np.zeros(0)
            self.time = 0.0
            self.cvel = np.zeros((cm.nbody, 6))
            self.xipos = np.zeros((cm.nbody, 3))
            self.sensordata = np.zeros(cm.nsensor)

        qpos = property(lambda s: s._qpos, lambda s, v: s._qpos.__setitem__(slice(None), v))
        qvel = property(lambda s: s._qvel, lambda s, v: s._qvel.__setitem__(slice(None), v))
        ctrl = property(lambda s: s._ctrl, lambda s, v: s._ctrl.__setitem__(slice(None), v))

    class mjtObj:
        mjOBJ_BODY = 1

    def mj_name2id(model, objtype, name):
        try:
            return cm.name2id(objtype, name)                    # the compiled model's name tables (mjOBJ_BODY = 1)
        except (ValueError, KeyError):
            return -1

    def mj_step(model, data, nstep=1):
        b = data._b
        b.set_state(data._qpos[None], data._qvel[None], None)   # picks up reset_idx's writes; the warm start survives (quirk Q3)
        b.physics_step(data._ctrl[None].astype(np.float64), int(nstep), 1)
        q, v, _ = b.get_state()
        data._qpos[:] = q[0]
        data._qvel[:] = v[0]
        data.cvel[:] = b.get(0, "cvel").reshape(cm.nbody, 6)
        data.xipos[:] = b.get(0, "xipos").reshape(cm.nbody, 3)
        data.sensordata[:] = b.get(0, "sensordata")
        data.time = float(b.get(0, "time")[0])

    def mju_negQuat(res, quat):
        res[0], res[1], res[2], res[3] = quat[0], -quat[1], -quat[2], -quat[3]

    def mju_rotVecQuat(res, vec, quat):
        w, x, y, z = quat
        vx, vy, vz = float(vec[0]), float(vec[1]), float(vec[2])
        # v' = v + 2 w (q x v) + 2 q x (q x v)  (unit quaternion)
        tx, ty, tz = 2 * (y * vz - z * vy), 2 * (z * vx - x * vz), 2 * (x * vy - y * vx)
        res[0] = vx + w * tx + (y * tz - z * ty)
        res[1] = vy + w * ty + (z * tx - x * tz)
        res[2] = vz + w * tz + (x * ty - y * tx)

    mj.MjModel, mj.MjData, mj.mjtObj = MjModel, MjData, mjtObj
    mj.mj_name2id, mj.mj_step, mj.mju_negQuat, mj.mju_rotVecQuat = mj_name2id, mj_step, mju_negQuat, mju_rotVecQuat
    viewer = types.ModuleType("mujoco.viewer")
    mj.viewer = viewer
    return mj, viewer


# ------------------------------------------------------------------------------------------------ the shared random source
class PhiloxCommands:
    """np.random.rand stand-in: the project's command uniforms (same definition as oracle/nm_oracle.c and the CUDA kernel)."""

    def __init__(self, env, seed):
        self.env, self.seed = env, seed
        self.phase, self.ids, self.draw = 0, None, 0

    def rand(self, *shape):
        step = int(self.env.common_step_counter)
        if self.ids is None and shape == (self.env.num_envs, 66):
            # observation noise (:300-301): uniform k of env i is word k % 4 of the Philox block with counter word 2 + k // 4
            out = np.zeros(shape)
            for i in range(shape[0]):
                for blk in range(17):
                    r = O.philox4x32(self.seed & 0xFFFFFFFF, i, step & 0xFFFFFFFF, step >> 32, 2 + blk, self.seed >> 32)
                    for q in range(4):
                        if 4 * blk + q < 66:
                            out[i, 4 * blk + q] = float(r[q] >> 8) / 16777216.0
            return out
        assert self.ids is not None and shape == (len(self.ids),), f"unexpected np.random.rand{shape} from the reference env"
        out = np.zeros(len(self.ids))
        for k, i in enumerate(self.ids):
            r = O.philox4x32(self.seed & 0xFFFFFFFF, int(i), step & 0xFFFFFFFF, step >> 32, self.phase, self.seed >> 32)
            out[k] = float(r[self.draw] >> 8) / 16777216.0
        self.draw += 1
        return out


def load_reference_env(ref, nmb_path=NMB):
    """The reference's NightmareV3Env class and config classes, imported from `ref` over the stub."""
    mj, viewer = make_mujoco_stub(nmb_path)
    saved = {k: sys.modules.get(k) for k in ("mujoco", "mujoco.viewer", "envs", "envs.nightmare_v3_env", "envs.nightmare_v3_config", "envs.helpers")}
    for k in ("envs", "envs.nightmare_v3_env", "envs.nightmare_v3_config", "envs.helpers"):
        sys.modules.pop(k, None)
    sys.modules["mujoco"], sys.modules["mujoco.viewer"] = mj, viewer
    # the reference's `envs` directory has no __init__.py (a namespace package), so a regular package of the same name on the
    # path -- this repository's import shim -- would win: pin the package to the reference directory explicitly
    pkg = types.ModuleType("envs")
    pkg.__path__ = [os.path.join(ref, "envs")]
    sys.modules["envs"] = pkg
    sys.path.insert(0, ref)
    try:
        import importlib
        mod = importlib.import_module("envs.nightmare_v3_env")
        cfgmod = importlib.import_module("envs.nightmare_v3_config")
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(ref)), mod.__file__
    finally:
        sys.path.remove(ref)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, cfgmod


def run_reference(ref, n, actions, ep0=None, seed=SEED, cfg_edit=None, via_reset=False):
    mod, cfgmod = load_reference_env(ref)
    cfg = cfgmod.NightmareV3Config()
    cfg.env.num_envs = n
    cfg.viewer.render = False
    cfg.viewer.record_states = False
    if cfg_edit:
        cfg_edit(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        env = mod.NightmareV3Env(cfg, log_dir="/tmp/nightmare_v3_refenv", num_threads=1)
    rng = PhiloxCommands(env, seed)
    real_rand = np.random.rand
    resample, reset_idx = env._resample_commands, env.reset_idx

    def resample_wrapped(env_ids):
        rng.ids, rng.draw = [int(i) for i in np.asarray(env_ids).reshape(-1)], 0
        resample(env_ids)
        rng.ids = None

    def reset_wrapped(env_ids):
        rng.phase = 1
        try:
            reset_idx(env_ids)
        finally:
            rng.phase = 0

    env._resample_commands, env.reset_idx = resample_wrapped, reset_wrapped
    np.random.rand = rng.rand
    rec = dict(obs=[], rew=[], done=[], time_out=[], commands=[], ep_len=[], qpos=[], ep_keys=None, ep_vals=[], ep_valid=[], time_outs_extra=[])
    try:
        if via_reset:                                            # the reference's reset(): reset_idx(all) + one zero-action step (:392-396)
            with contextlib.redirect_stdout(io.StringIO()):
                obs0, priv0 = env.reset()
            assert priv0 is None
            rec["reset_obs"] = obs0.numpy().copy()
        else:
            env.reset_idx(np.arange(n))                          # what the GPU class and the oracle do before the first step
        if ep0 is not None:
            env.episode_length_buf = torch.tensor(np.array(ep0), dtype=torch.int64)   # a copy, re-bound like rsl_rl does (train.py:54)
        for a in actions:
            obs, priv, rew, done, extras = env.step(torch.from_numpy(np.ascontiguousarray(a)))
            assert priv is None
            rec["obs"].append(obs.numpy().copy()); rec["rew"].append(rew.numpy().copy()); rec["done"].append(done.numpy().copy())
            rec["time_out"].append(np.asarray(env.time_out_buf).copy())
            rec["commands"].append(env.commands.astype(np.float32).copy())
            rec["ep_len"].append(np.asarray(env.episode_length_buf).astype(np.int64).copy())
            rec["qpos"].append(np.array([d.qpos.copy() for d in env.data], dtype=np.float32))
            ep = extras.get("episode")
            if ep:
                keys = sorted(ep.keys())
                rec["ep_keys"] = rec["ep_keys"] or keys
                rec["ep_vals"].append(np.array([float(ep[k]) for k in keys], dtype=np.float32)); rec["ep_valid"].append(True)
            else:
                rec["ep_vals"].append(np.zeros(len(rec["ep_keys"] or []) or 8, dtype=np.float32)); rec["ep_valid"].append(False)
            to = extras.get("time_outs")
            rec["time_outs_extra"].append(to.numpy().copy() if to is not None else np.zeros(n, dtype=np.float32))
    finally:
        np.random.rand = real_rand
    out = {k: np.array(v) for k, v in rec.items() if k != "ep_keys"}
    out["ep_keys"] = np.array(rec["ep_keys"] or [])
    return out


def scenarios():
    """(name, n, actions[T,n,18], ep0, cfg_edit).  Chosen to hit the quirks: resampling at 625, time-out at 1251, falls, stale buffers."""
    rng = np.random.default_rng(SEED)
    n, T = 16, 80
    a = rng.normal(size=(T, n, 18)).astype(np.float32)
    a[:, 3] *= 8.0                                              # env 3 thrashes: contact-force / tilt terminations
    a[:, 7] *= 8.0
    ep0 = np.array([0, 100, 620, 621, 1240, 1245, 1249, 1250, 5, 50, 500, 623, 624, 1100, 1200, 1248], dtype=np.int64)
    yield "default", n, a, ep0, None

    def strict(cfg):                                            # the two optional termination clauses and every inactive reward term on
        cfg.env.tibia_contact_mode = 2
        cfg.env.body_contact_mode = 2
        s = cfg.rewards.scales
        s.lin_vel_z, s.ang_vel_xy, s.base_height, s.torques, s.dof_vel, s.feet_air_time, s.stand_still, s.feet_contact_forces = (
            -2.0, -0.05, -1.0, -1e-5, -1e-4, 1.0, -0.5, -0.01)
    yield "all_terms", n, a[:50], ep0, strict

    def noisy(cfg):                                             # observation noise with the reference's 12-DoF slice boundaries (quirk Q8)
        cfg.noise.add_noise = True
    yield "noise", n, a[:20], ep0, noisy
    yield "via_reset", n, a[:10], None, None                    # starts with the reference's reset() instead of reset_idx(all)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {}
    for name, n, a, ep0, edit in scenarios():
        r = run_reference(ref, n, a, ep0, cfg_edit=edit, via_reset=(name == "via_reset"))
        out[f"{name}.actions"] = a
        out[f"{name}.ep0"] = ep0 if ep0 is not None else np.zeros(0, dtype=np.int64)
        for k, v in r.items():
            out[f"{name}.{k}"] = v
        print(f"{name}: {len(a)} steps x {n} envs, resets {int(r['done'].sum())}, time-outs {int(r['time_out'].sum())}, "
              f"episode extras on {int(r['ep_valid'].sum())} steps, keys {list(r['ep_keys'])}")
    path = os.path.join(ROOT, "tests", "golden", "reference_env_on_oracle_physics.npz")
    np.savez_compressed(path, seed=SEED, **out)
    print("wrote", path, f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
