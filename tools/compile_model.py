#!/usr/bin/env python
"""Compile an MJCF scene into the committed model asset used at run time.

    python tools/compile_model.py /root/reference/Code/mujoco/our_robot/walking_scene.xml \
        opendog_b200/assets/our_robot.model.json

The GPU box has no /root/reference, so bench/tests/smoke load the committed JSON; this script
(and tests/test_model_compile.py, which re-derives the asset when the reference is mounted) is
how it was made.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendog_b200.model.compile import compile_model, save_compiled  # noqa: E402
from opendog_b200.model.mjcf import load_mjcf  # noqa: E402


def main():
    src, dst = sys.argv[1], sys.argv[2]
    mode = sys.argv[3] if len(sys.argv) > 3 else "legacy"
    d = compile_model(load_mjcf(src, inertia_mode=mode))
    d["inertia_mode"] = mode
    save_compiled(d, dst)
    print(f"{dst}: nleg={d['nleg']} njl={d['njl']} nq={d['nq']} nv={d['nv']} nu={d['nu']} "
          f"ngeom={d['ngeom']} nvert={d['nvert']}")


if __name__ == "__main__":
    main()
