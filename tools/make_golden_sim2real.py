#!/usr/bin/env python
"""Generate tests/golden/sim2real_env_golden.npz by running the REFERENCE's own `QuadrupedEnv`
(/root/reference/Code/mujoco/sim2real/train.py:151-411, imported unmodified) with the third-party `mujoco`
module stubbed onto the oracle physics (oracle/odg_oracle.c). Needs /root/reference (absent on the GPU box).

Pinned: the env logic the reference itself implements — 4->8 symmetric-trot action mapping and ctrlrange clipping,
the 22-float observation, the nine reward terms, the three termination rules, reset-with-settle — on a given
physics state sequence. Not pinned: the physics (PARITY UNPINNED, see oracle/odg_oracle.h).
The script also asserts that oracle/sim2real_oracle.py reproduces every recorded value.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/mujoco"
sys.path.insert(0, ROOT)

from oracle.oracle import Sim  # noqa: E402
from oracle.sim2real_oracle import QuadrupedEnvOracle  # noqa: E402


def install_mujoco_stub():
    mj = types.ModuleType("mujoco")

    class FatalError(Exception):
        pass

    class _Obj:
        mjOBJ_ACTUATOR, mjOBJ_JOINT, mjOBJ_KEY = 1, 2, 3

    class _Opt:
        pass

    class MjModel:
        def __init__(self):
            self.sim_desc = Sim("our_robot").desc
            d = self.sim_desc
            self.opt = _Opt(); self.opt.timestep = d["timestep"]
            self.actuator_ctrlrange = np.array(d["act_ctrlrange"])
            self.act_names = list(d["act_names"])
            self.joint_names = ["root"] + [n for leg in d["joint_names"] for n in leg]       # joint 0 = free joint
            self.actuator_trnid = np.array([[1 + d["act_leg"][u] * d["njl"] + d["act_joint"][u], -1] for u in range(d["nu"])])
            self.jnt_qposadr = np.array([0] + [7 + i for i in range(d["nv"] - 6)])
            self.jnt_dofadr = np.array([0] + [6 + i for i in range(d["nv"] - 6)])

        @staticmethod
        def from_xml_path(path):
            assert path.endswith("walking_scene.xml")
            return MjModel()

    class MjData:
        def __init__(self, model):
            self._sim = Sim("our_robot")
            self.qpos, self.qvel, self.ctrl = self._sim.qpos, self._sim.qvel, self._sim.ctrl

    def mj_name2id(model, objtype, name):
        if objtype == _Obj.mjOBJ_ACTUATOR:
            return model.act_names.index(name) if name in model.act_names else -1
        if objtype == _Obj.mjOBJ_JOINT:
            return model.joint_names.index(name) if name in model.joint_names else -1
        if objtype == _Obj.mjOBJ_KEY:
            return 0 if name == "home" else -1
        return -1

    def mj_id2name(model, objtype, i):
        return model.joint_names[i] if objtype == _Obj.mjOBJ_JOINT else model.act_names[i]

    def mj_resetDataKeyframe(model, data, key):
        data._sim.reset_keyframe()

    def mj_forward(model, data):
        data._sim.forward()

    def mj_step(model, data):
        data._sim.step()

    mj.FatalError, mj.mjtObj, mj.MjModel, mj.MjData = FatalError, _Obj, MjModel, MjData
    mj.mj_name2id, mj.mj_id2name, mj.mj_resetDataKeyframe, mj.mj_forward, mj.mj_step = \
        mj_name2id, mj_id2name, mj_resetDataKeyframe, mj_forward, mj_step
    viewer = types.ModuleType("mujoco.viewer")
    mj.viewer = viewer
    sys.modules["mujoco"] = mj
    sys.modules["mujoco.viewer"] = viewer


def main():
    install_mujoco_stub()
    sys.path.insert(0, os.path.join(REF, "sim2real"))
    import train as ref                                               # the reference module, unmodified
    # the global maps train() fills in (train.py:512-523), replayed verbatim on the stub model
    import mujoco
    vm = mujoco.MjModel.from_xml_path("our_robot/walking_scene.xml"); vd = mujoco.MjData(vm)
    for act_name in ref.ACTUATOR_NAMES_ORDERED:
        act_id = mujoco.mj_name2id(vm, mujoco.mjtObj.mjOBJ_ACTUATOR, act_name)
        j_id = vm.actuator_trnid[act_id, 0]; j_name = mujoco.mj_id2name(vm, mujoco.mjtObj.mjOBJ_JOINT, j_id)
        ref.ACTUATOR_TO_JOINT_NAME_MAP[act_name] = j_name; ref.JOINT_NAME_TO_QPOS_IDX_MAP[j_name] = vm.jnt_qposadr[j_id]
    mujoco.mj_resetDataKeyframe(vm, vd, 0)
    for act_name in ref.ACTUATOR_NAMES_ORDERED:
        j_name = ref.ACTUATOR_TO_JOINT_NAME_MAP[act_name]
        ref.sim_keyframe_home_qpos_map[j_name] = vd.qpos[ref.JOINT_NAME_TO_QPOS_IDX_MAP[j_name]]

    import builtins
    real_print = builtins.print
    builtins.print = lambda *a, **k: None                             # the reference prints a debug line per step
    env = ref.QuadrupedEnv("our_robot/walking_scene.xml")
    orc = QuadrupedEnvOracle()
    rng = np.random.default_rng(7)
    n_ep, n_steps = 3, 40
    rec = {k: [] for k in ("action", "obs", "reward", "done", "reason", "sim_target_rad", "reset_obs", "inject")}
    reasons = ["max_steps", "mj_error", "orientation_limit", "too_much_backward"]
    for ep in range(n_ep):
        ro = env.reset(); oo = orc.reset()
        assert np.array_equal(ro, oo)
        per = {k: [] for k in rec}
        per["reset_obs"] = ro
        for t in range(n_steps):
            a = rng.uniform(-1, 1, 4).astype(np.float32)
            if ep == 1:
                a[0] = a[2] = -abs(a[0])                              # push backwards: "too_much_backward"
            inj = np.full(4, np.nan)
            if ep == 2 and t == 3:                                   # roll the trunk 69 degrees: "orientation_limit"
                inj[:] = [np.cos(0.6), np.sin(0.6), 0.0, 0.0]
                env.data.qpos[3:7] = inj; orc.sim.qpos[3:7] = inj
            # NumPy: the reference ran NumPy 1.26 (float32 scalar * Python float -> float64); under NumPy >= 2 (NEP 50)
            # the same expression stays float32. Passing the float32 action widened to float64 reproduces the 1.26 values.
            robs, rr, rd, rinfo = env.step(a.astype(np.float64))
            oobs, orr, od, oinfo = orc.step(a)
            assert np.array_equal(robs, oobs), (ep, t)
            assert abs(rr - orr) <= 1e-12 * max(1.0, abs(rr)), (ep, t, rr, orr)
            assert rd == od and rinfo["termination_reason"] == oinfo["termination_reason"]
            assert np.array_equal(rinfo["sim_target_rad"], oinfo["sim_target_rad"])
            per["action"].append(a); per["obs"].append(robs); per["reward"].append(rr); per["done"].append(rd)
            per["reason"].append(reasons.index(rinfo["termination_reason"])); per["sim_target_rad"].append(rinfo["sim_target_rad"])
            per["inject"].append(inj)
            if rd:
                break
        n = len(per["action"])
        pad = n_steps - n
        for k in ("action", "obs", "reward", "done", "reason", "sim_target_rad", "inject"):
            arr = np.array(per[k])
            if pad:
                arr = np.concatenate([arr, np.zeros((pad,) + arr.shape[1:], arr.dtype)])
            rec[k].append(arr)
        rec["reset_obs"].append(per["reset_obs"])
        rec.setdefault("length", []).append(n)
    builtins.print = real_print
    out = {k: np.stack(v) if k != "length" else np.array(v) for k, v in rec.items()}
    path = os.path.join(ROOT, "tests", "golden", "sim2real_env_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: lengths {out['length']}, reasons at end {[int(out['reason'][e, out['length'][e] - 1]) for e in range(n_ep)]}")


if __name__ == "__main__":
    main()
