"""The physics oracle against REAL MuJoCo trajectories — skipped until someone runs tools/make_golden_mujoco.py on a
machine that has `mujoco==3.2.3` and commits tests/golden/mj_<model>.npz. No such machine was available to this project
(profiles/r02_mujoco_probe.txt), which is why oracle/ says PARITY UNPINNED at the single-step level for the physics (the shipped walk files pin it
loosely at the trajectory level: tests/test_mujoco_pin_walk_json.py); this file is the gate that
pins it the moment the fixture exists. Tolerances are the ones an exact restatement must meet in double precision with
the oracle's tight solver settings; a failure here is a finding about the restatement (mesh inertia, invweight0,
PlaneConvex support points, ...), not about the CUDA path."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["our_robot", "go1"])
def test_oracle_follows_mujoco_step_by_step(name):
    path = os.path.join(GOLD, f"mj_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("no MuJoCo fixture: run tools/make_golden_mujoco.py where mujoco==3.2.3 is installed")
    from oracle.oracle import Sim
    g = np.load(path)
    s = Sim(name)
    # model constants the compiler derived without MuJoCo
    nb = 1 + s.desc["nleg"] * s.desc["njl"]
    mass = np.r_[s.desc["base_mass"], np.array(s.desc["mass"])[:, :s.desc["njl"]].reshape(-1)]
    mj_mass = g["model_body_mass"][1:]
    if len(mj_mass) == nb:                                   # (OpenDOG's welded paws are fused into the calves here)
        assert np.allclose(mass, mj_mass, rtol=1e-9)
    assert np.allclose(np.array(s.desc["key_qpos"]), g["model_key_qpos"][0], atol=1e-12)
    inv = np.r_[s.desc["base_invweight0"], np.array(s.desc["dof_invweight0"])[:, :s.desc["njl"]].reshape(-1)]
    assert np.allclose(inv, g["model_dof_invweight0"], rtol=1e-6), np.abs(inv / g["model_dof_invweight0"] - 1).max()
    # single steps from identical states
    worst_q = worst_v = 0.0
    n = len(g["ncon"])
    for i in range(0, n, 7):
        s.reset_keyframe()
        s.qpos[:] = g["qpos0"][i]; s.qvel[:] = g["qvel0"][i]; s.qacc_warmstart[:] = g["warm0"][i]; s.ctrl[:] = g["ctrl"][i]
        s.step()
        worst_q = max(worst_q, np.abs(s.qpos - g["qpos1"][i]).max())
        worst_v = max(worst_v, np.abs(s.qvel - g["qvel1"][i]).max())
        assert s.ncon == g["ncon"][i], (i, s.ncon, int(g["ncon"][i]))
    assert worst_q < 1e-7 and worst_v < 1e-5, (worst_q, worst_v)
