#!/usr/bin/env python
"""Cross-check of the CPU oracle (and, through it, of the CUDA step) against REAL MuJoCo.

STATUS: NOT RUN.  `mujoco` (pinned `<=3.1.2` by the reference, requirements.txt:2) is not installable in the build
environment (no wheel, no network), so nothing in this repository has been compared with MuJoCo itself; the oracle's
parity is "unpinned" (DESIGN.md §2).  This script is what a maintainer runs on a machine that has `pip install
mujoco==3.1.2` and the reference's model directory:

    python tools/mujoco_crosscheck.py --model /path/to/nightmare_rl/models/nightmare_v3/mjmodel.xml [--steps 200]

It (1) compiles the MJCF with our compiler and diffs the model constants against `MjModel` (masses, inertias, body /
geom frames, hull sizes, invweight0, actuator parameters, options), (2) replays the BASELINE.json configs[0] action
sequence (U(-1,1) joint targets from torch seed 0 through the env's PD law) in both engines from `qpos0`, one `mj_step` at
a time with the oracle re-synchronised to MuJoCo's state before every substep, and reports per-stage maximum deviations
(xpos, cinert, M, qfrc_bias, qacc_smooth, contact count / geoms / dist, efc_force, sensordata, qpos/qvel after the step).
Each item marked "❓ recalled" in SURVEY.md Appendix A shows up here as a stage whose deviation is not ~1e-12.

Which MODEL OPTION repairs which mismatch (all are slots of the .nmb file, written by nightmare_rl_b200/mjcf.py and honoured
by oracle and kernel -- change the constant there or the array in the file, re-save, re-run; no engine code changes):

    stage that deviates                         option to try                                        slot
    ------------------------------------------  ---------------------------------------------------  ------------
    ncon: more / fewer contacts on a flat hull  PLANEMESH_MAXCON (contacts per plane-mesh pair)       opt_int[7]
    ncon / contact vertices: other vertices     PLANEMESH_ALLVERTS (neighbours vs all hull vertices)  opt_int[9]
    ncon: a close second contact kept/dropped   PLANEMESH_SEP (fraction of rbound), PLANEMESH_SEPVERT opt_real[9], opt_int[10]
    efc_R / efc_D of contact rows off by const  PYRAMID_RFAC (R = fac * mu_reg^2 * R[first])          opt_real[10]
    qacc_warmstart after the step               WARM_AFTER_NOSLIP (save point relative to noslip)     opt_int[11]
    tibia-tibia penetration depth / iterations  MPR_ITERATIONS, MPR_TOLERANCE                         opt_int[8], opt_real[8]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True, help="mjmodel.xml of the reference (models/nightmare_v3/mjmodel.xml)")
    ap.add_argument("--steps", type=int, default=200, help="env steps (2 substeps each) to replay")
    ap.add_argument("--tol", type=float, default=1e-9)
    ap.add_argument("--write-golden", action="store_true",
                    help="also record MuJoCo's own trajectory for the BASELINE configs[0] action sequence as tests/golden/mujoco_config1.npz; "
                         "tests/test_mujoco_golden.py (skipped while that file is absent) then pins the oracle's physics against it")
    a = ap.parse_args()
    try:
        import mujoco as mj
    except ImportError:
        print("mujoco is not importable here: the cross-check cannot run (this is the expected outcome in the build sandbox).")
        return 2
    import tempfile

    import torch

    from nightmare_rl_b200 import mjcf
    from oracle import oracle as O

    cm = mjcf.compile_mjcf(a.model)
    m = mj.MjModel.from_xml_path(a.model)
    d = mj.MjData(m)
    print(f"MuJoCo {mj.__version__}; nq/nv/nu/nbody/ngeom/nsensor: mujoco {m.nq}/{m.nv}/{m.nu}/{m.nbody}/{m.ngeom}/{m.nsensor} "
          f"ours {cm.nq}/{cm.nv}/{cm.nu}/{cm.nbody}/{cm.ngeom}/{cm.nsensor}")

    def diff(name, ours, theirs):
        ours, theirs = np.asarray(ours, dtype=np.float64).ravel(), np.asarray(theirs, dtype=np.float64).ravel()
        if ours.shape != theirs.shape:
            print(f"  {name:28s} SHAPE {ours.shape} vs {theirs.shape}")
            return np.inf
        e = float(np.abs(ours - theirs).max()) if ours.size else 0.0
        print(f"  {name:28s} max |diff| {e:.3e}" + ("" if e <= a.tol * max(1.0, float(np.abs(theirs).max()) if theirs.size else 1.0) else "   <-- DIFFERS"))
        return e

    A = cm.arrays
    print("model constants:")
    for ours, theirs in (("body_mass", m.body_mass), ("body_inertia", m.body_inertia), ("body_pos", m.body_pos), ("body_quat", m.body_quat),
                         ("body_ipos", m.body_ipos), ("body_iquat", m.body_iquat), ("body_invweight0", m.body_invweight0),
                         ("dof_invweight0", m.dof_invweight0), ("dof_armature", m.dof_armature), ("dof_damping", m.dof_damping),
                         ("jnt_axis", m.jnt_axis), ("jnt_pos", m.jnt_pos), ("qpos0", m.qpos0), ("geom_pos", m.geom_pos), ("geom_quat", m.geom_quat),
                         ("geom_rbound", m.geom_rbound), ("geom_friction", m.geom_friction), ("geom_solref", m.geom_solref),
                         ("geom_solimp", m.geom_solimp), ("site_pos", m.site_pos), ("act_gear", m.actuator_gear[:, 0]),
                         ("act_ctrlrange", m.actuator_ctrlrange)):
        if ours in A:
            diff(ours, A[ours], theirs)
    print(f"  total mass: ours {A['body_mass'].sum():.6f}  mujoco {m.body_mass.sum():.6f}")
    for g in range(m.ngeom):
        if m.geom_type[g] == mj.mjtGeom.mjGEOM_MESH:
            mid = m.geom_dataid[g]
            print(f"  geom {g}: hull vertices ours {A['geom_hull_num'][g]}  mujoco mesh graph "
                  f"{'present' if m.mesh_graphadr[mid] >= 0 else 'absent'} (mesh verts {m.mesh_vertnum[mid]})")

    with tempfile.TemporaryDirectory() as td:
        nmb = os.path.join(td, "m.nmb")
        cm.save(nmb)
        om = O.OracleModel(nmb)
        ob = O.OracleBatch(om, 1)
        gen = torch.Generator().manual_seed(0)
        default = np.array([0.0, np.pi / 5, 0.0] * 6)
        worst = {}
        for t in range(a.steps):
            act = np.clip(torch.rand(18, generator=gen).numpy() * 2 - 1, -1, 1) * 0.2
            ctrl = ((act - default) - d.qpos[-18:]) * 20.0
            d.ctrl[:] = ctrl
            for sub in range(2):
                ob.set_state(d.qpos[None].copy(), d.qvel[None].copy(), d.qacc_warmstart[None].copy())
                mj.mj_forward(m, d)                          # every stage of the substep, without integrating
                ob.forward(ctrl[None].astype(np.float64))
                stages = {"xpos": d.xpos, "xipos": d.xipos, "subtree_com": d.subtree_com, "cinert": d.cinert, "cdof": d.cdof,
                          "cvel": d.cvel, "qfrc_bias": d.qfrc_bias, "qfrc_actuator": d.qfrc_actuator, "qacc_smooth": d.qacc_smooth,
                          "qfrc_constraint": d.qfrc_constraint, "qacc": d.qacc, "sensordata": d.sensordata}
                for k, v in stages.items():
                    e = float(np.abs(ob.get(0, k).reshape(-1) - np.asarray(v).reshape(-1)).max())
                    worst[k] = max(worst.get(k, 0.0), e)
                Mfull = np.zeros((m.nv, m.nv))
                mj.mj_fullM(m, Mfull, d.qM)
                worst["M"] = max(worst.get("M", 0.0), float(np.abs(ob.get(0, "M").reshape(m.nv, m.nv) - Mfull).max()))
                ncon_o = int(ob.get(0, "ncon")[0])
                worst["ncon mismatch steps"] = worst.get("ncon mismatch steps", 0) + int(ncon_o != d.ncon)
                if ncon_o == d.ncon and d.ncon:
                    con = ob.get(0, "contact").reshape(-1, 7)
                    worst["contact dist"] = max(worst.get("contact dist", 0.0), float(np.abs(con[:, 3] - d.contact.dist[: d.ncon]).max()))
                    worst["contact pos"] = max(worst.get("contact pos", 0.0), float(np.abs(con[:, 4:7] - d.contact.pos[: d.ncon]).max()))
                    worst["contact geom ids differ"] = worst.get("contact geom ids differ", 0) + int(
                        (con[:, 0].astype(int) != d.contact.geom1[: d.ncon]).any() or (con[:, 1].astype(int) != d.contact.geom2[: d.ncon]).any())
                    if d.nefc == int(ob.get(0, "nefc")[0]):
                        worst["efc_force"] = max(worst.get("efc_force", 0.0), float(np.abs(ob.get(0, "efc_force") - d.efc_force).max()))
                mj.mj_step(m, d)
                ob.physics_step(ctrl[None].astype(np.float64), 1, 1)
                q, v, w = ob.get_state()
                worst["qpos after step"] = max(worst.get("qpos after step", 0.0), float(np.abs(q[0] - d.qpos).max()))
                worst["qvel after step"] = max(worst.get("qvel after step", 0.0), float(np.abs(v[0] - d.qvel).max()))
                worst["qacc_warmstart"] = max(worst.get("qacc_warmstart", 0.0), float(np.abs(w[0] - d.qacc_warmstart).max()))
        if a.write_golden:
            d2 = mj.MjData(m)
            gen2 = torch.Generator().manual_seed(0)
            rec = dict(qpos=[], qvel=[], warm=[], ctrl=[], sensordata=[], ncon=[])
            for t in range(a.steps):
                act = np.clip(torch.rand(18, generator=gen2).numpy() * 2 - 1, -1, 1) * 0.2
                c = ((act - default) - d2.qpos[-18:]) * 20.0
                d2.ctrl[:] = c
                for sub in range(2):
                    rec["ctrl"].append(c.copy())
                    mj.mj_step(m, d2)
                    rec["qpos"].append(d2.qpos.copy()); rec["qvel"].append(d2.qvel.copy()); rec["warm"].append(d2.qacc_warmstart.copy())
                    rec["sensordata"].append(d2.sensordata.copy()); rec["ncon"].append(d2.ncon)
            out = os.path.join(ROOT, "tests", "golden", "mujoco_config1.npz")
            np.savez_compressed(out, version=mj.__version__, **{k: np.array(v) for k, v in rec.items()})
            print("wrote", out)
        print(f"per-stage worst deviation over {a.steps} env steps (oracle re-synchronised to MuJoCo before every substep):")
        for k, v in worst.items():
            print(f"  {k:28s} {v:.3e}" if isinstance(v, float) else f"  {k:28s} {v}")
        print("a stage that is not ~1e-12: see the option table in this file's docstring (model data, not code)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
