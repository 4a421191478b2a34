#!/usr/bin/env python
"""tests/golden/trot_golden.npz: the reference's hand-coded trot (`create_control_sequence`,
/root/reference/Code/mujoco/sim2real/main.py:63-151 — canonical input sequence (i) of SURVEY section 8c) evaluated by
the reference's OWN function source. main.py cannot be imported (its top level loads a model and opens a viewer), so
this script parses it and executes only its imports-free pieces unmodified: the two function definitions
(`convert_sim_rad_to_real_deg`, `create_control_sequence`) and the three configuration tables they are called with
(`ACTUATOR_NAMES`, `real_robot_home_deg_map`, `joint_scale_factors`). The model arguments are what main.py:184-228
reads from MuJoCo: the actuator ids, `model.nu`, `model.actuator_ctrlrange` and the ctrl of keyframe 'home' — taken
from the compiled model (opendog_b200/assets/our_robot.model.json, itself derived from our_robot.xml)."""
import ast
import builtins
import json
import math
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Code/mujoco/sim2real/main.py"


def reference_namespace():
    tree = ast.parse(open(REF).read())
    keep_fn = {"convert_sim_rad_to_real_deg", "create_control_sequence"}
    keep_var = {"ACTUATOR_NAMES", "real_robot_home_deg_map", "joint_scale_factors"}
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in keep_fn:
            body.append(node)
        elif isinstance(node, ast.Assign) and len(node.targets) == 1 and getattr(node.targets[0], "id", None) in keep_var:
            body.append(node)
    assert len(body) == 5, [getattr(n, "name", None) for n in body]
    ns = {"np": np, "math": math, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    from opendog_b200.model.compile import load_compiled
    d = load_compiled("our_robot")
    ns = reference_namespace()
    names = ns["ACTUATOR_NAMES"]
    act_id = {n: d["act_names"].index(n) for n in names}                       # mj_name2id(model, mjOBJ_ACTUATOR, name)
    model = types.SimpleNamespace(nu=d["nu"], actuator_ctrlrange=np.array(d["act_ctrlrange"], dtype=np.float64))
    sim_home = {n: float(d["key_ctrl"][act_id[n]]) for n in names}             # data.ctrl after mj_resetDataKeyframe('home')
    seq = ns["create_control_sequence"](sim_home, ns["real_robot_home_deg_map"], ns["joint_scale_factors"], act_id, model)
    S = len(seq)
    rad = np.zeros((S, d["nu"])); deg = np.zeros((S, d["nu"])); dur = np.zeros(S)
    for i, (sim_rad, real_deg, duration) in enumerate(seq):
        for n in names:
            rad[i, act_id[n]] = sim_rad[n]; deg[i, act_id[n]] = real_deg[n]
        dur[i] = duration
    # the JSON main.py:236-243 writes (rounded)
    js = [{"duration": round(du, 3), "targets_deg": {n: round(float(real_deg[n]), 2) for n in real_deg}} for _, real_deg, du in seq]
    path = os.path.join(ROOT, "tests", "golden", "trot_golden.npz")
    np.savez_compressed(path, targets_rad=rad, targets_deg=deg, durations=dur, json=np.array(json.dumps(js)))
    print("wrote", path, rad.shape, "total", dur.sum(), "s")


if __name__ == "__main__":
    main()
