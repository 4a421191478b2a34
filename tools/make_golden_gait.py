#!/usr/bin/env python
"""tests/golden/gait_golden.npz: the reference's own `load_and_process_sequence` + angle-map functions
(/root/reference/Code/mujoco/sim2real/run.py:60-79,176-240 and train.py:120-130, imported unmodified with the
`mujoco` module stubbed) applied to the gait files that ship in the reference (sim2real/walk.json,
Code/examples/walks/forward1.json). The JSON inputs are stored in the fixture so the test needs no reference."""
import builtins
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_sim2real import install_mujoco_stub  # noqa: E402

REF = "/root/reference/Code"


def main():
    install_mujoco_stub()
    sys.path.insert(0, os.path.join(REF, "mujoco", "sim2real"))
    real_print = builtins.print
    builtins.print = lambda *a, **k: None
    import run as ref_run
    import train as ref_train
    import mujoco
    model = mujoco.MjModel.from_xml_path("our_robot/walking_scene.xml")
    model.nu = 8
    act_id_map = {n: model.act_names.index(n) for n in ref_run.ACTUATOR_NAMES}
    desc = model.sim_desc
    sim_home = {n: desc["key_ctrl"][act_id_map[n]] for n in ref_run.ACTUATOR_NAMES}
    out = {}
    for tag, path in (("walk", os.path.join(REF, "mujoco/sim2real/walk.json")), ("forward1", os.path.join(REF, "examples/walks/forward1.json"))):
        seq = ref_run.load_and_process_sequence(path, sim_home, ref_run.REAL_ROBOT_HOME_DEG_MAP, ref_run.JOINT_SCALE_FACTORS,
                                                act_id_map, list(ref_run.ACTUATOR_NAMES), model)
        tg = np.full((len(seq), 8), np.nan); du = np.zeros(len(seq))
        for i, (sim_targets, dur, _) in enumerate(seq):
            for n, v in sim_targets.items():
                tg[i, act_id_map[n]] = v
            du[i] = dur
        out[tag + "_targets"] = tg; out[tag + "_durations"] = du
        out[tag + "_json"] = np.array(json.dumps(json.load(open(path))))
    # exporter's angle map (train.py:120-130) on a grid of commanded sim angles
    ref_train.ACTUATOR_TO_JOINT_NAME_MAP.update({n: n.replace("_actuator", "_joint") for n in ref_train.ACTUATOR_NAMES_ORDERED})
    home_q = {n.replace("_actuator", "_joint"): desc["key_qpos"][7 + desc["act_leg"][act_id_map[n]] * 2 + desc["act_joint"][act_id_map[n]]]
              for n in ref_train.ACTUATOR_NAMES_ORDERED}
    grid = np.linspace(-2.0, 3.0, 11)
    conv = np.array([[ref_train.convert_sim_rad_to_real_deg(n, float(x), home_q, ref_train.real_robot_home_deg_map,
                                                            ref_train.joint_scale_factors) for x in grid]
                     for n in ref_train.ACTUATOR_NAMES_ORDERED])
    out["export_grid"] = grid; out["export_deg"] = conv
    builtins.print = real_print
    path = os.path.join(ROOT, "tests", "golden", "gait_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
