#!/usr/bin/env python
"""Pin the physics oracle to REAL MuJoCo — for a machine that has `mujoco==3.2.3` (Code/mujoco/install.sh:27).

    python tools/make_golden_mujoco.py /path/to/OpenDOG/Code/mujoco [our_robot|go1]

NOT RUN IN THIS PROJECT'S CONTAINERS: neither the build container nor the GPU boxes have the wheel, an index or a
libmujoco (profiles/r02_mujoco_probe.txt), so oracle/ is "parity unpinned" for the physics. This script is the missing
step, written against the public Python bindings: it writes tests/golden/mj_<model>.npz with
  * the mjModel constants the model compiler derives on its own (masses, inertial frames, mesh-derived inertias,
    invweight0, mixed contact parameters' ingredients, keyframe), for tests/test_golden_mujoco.py to diff against
    opendog_b200/assets/<model>.model.json, and
  * seeded trajectories in the form SURVEY section 8d config 1 asks for — keyframe reset, a fresh ctrl ~ U(ctrlrange)
    every step, 1000 x mj_step — with the full state BEFORE every step (qpos, qvel, qacc_warmstart, ctrl) and qpos / qvel /
    ncon / contact dist / efc_force AFTER it, so the oracle can be compared step by step from identical states.
tests/test_golden_mujoco.py is skipped until such a file exists.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = {"our_robot": "our_robot/walking_scene.xml", "go1": "unitree_go1/walk_scene.xml"}


def main():
    import mujoco                                        # the reference's own dependency; absent here
    ref, name = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "our_robot")
    m = mujoco.MjModel.from_xml_path(os.path.join(ref, SCENES[name]))
    d = mujoco.MjData(m)
    out = {"mujoco_version": np.array(mujoco.__version__)}
    for k in ("body_mass", "body_ipos", "body_iquat", "body_inertia", "body_pos", "body_quat", "body_invweight0",
              "dof_invweight0", "dof_armature", "dof_frictionloss", "dof_damping", "jnt_pos", "jnt_axis", "jnt_range",
              "geom_type", "geom_bodyid", "geom_size", "geom_pos", "geom_quat", "geom_friction", "geom_margin", "geom_gap",
              "geom_solref", "geom_solimp", "geom_solmix", "geom_condim", "geom_priority", "geom_contype", "geom_conaffinity",
              "actuator_ctrlrange", "actuator_forcerange", "actuator_gainprm", "actuator_biasprm", "key_qpos", "key_ctrl",
              "mesh_vert", "mesh_vertadr", "mesh_vertnum", "mesh_graph", "mesh_graphadr", "geom_dataid"):
        out["model_" + k] = np.array(getattr(m, k))
    out["opt"] = np.array([m.opt.timestep, m.opt.impratio, float(m.opt.cone), *m.opt.gravity, m.opt.tolerance,
                           float(m.opt.iterations), m.opt.ls_tolerance, float(m.opt.ls_iterations)])
    rng = np.random.default_rng(0)
    lo, hi = m.actuator_ctrlrange[:, 0], m.actuator_ctrlrange[:, 1]
    rec = {k: [] for k in ("qpos0", "qvel0", "warm0", "ctrl", "qpos1", "qvel1", "qacc", "ncon", "fn_sum", "contact_dist",
                           "contact_geom", "efc_force_sum")}
    for traj in range(4):
        mujoco.mj_resetDataKeyframe(m, d, 0)
        if traj >= 2:                                    # tilted drops: contacts on more than the feet
            d.qpos[3:7] = [np.cos(0.3 * traj), np.sin(0.3 * traj), 0, 0]
            d.qpos[2] += 0.1
        for t in range(1000):
            d.ctrl[:] = rng.uniform(lo, hi)
            rec["qpos0"].append(d.qpos.copy()); rec["qvel0"].append(d.qvel.copy()); rec["warm0"].append(d.qacc_warmstart.copy())
            rec["ctrl"].append(d.ctrl.copy())
            mujoco.mj_step(m, d)
            rec["qpos1"].append(d.qpos.copy()); rec["qvel1"].append(d.qvel.copy()); rec["qacc"].append(d.qacc.copy())
            rec["ncon"].append(d.ncon)
            dist = np.full(48, np.nan); geom = np.full(48, -1); fn = 0.0
            f6 = np.zeros(6)
            for c in range(min(d.ncon, 48)):
                dist[c] = d.contact[c].dist; geom[c] = d.contact[c].geom2
                mujoco.mj_contactForce(m, d, c, f6); fn += f6[0]
            rec["contact_dist"].append(dist); rec["contact_geom"].append(geom); rec["fn_sum"].append(fn)
            rec["efc_force_sum"].append(float(np.sum(d.efc_force)) if d.nefc else 0.0)
    out.update({k: np.array(v) for k, v in rec.items()})
    path = os.path.join(ROOT, "tests", "golden", f"mj_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(rec['ncon'])} steps, MuJoCo {mujoco.__version__}")


if __name__ == "__main__":
    main()
