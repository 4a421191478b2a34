#!/usr/bin/env python
"""Generate tests/golden/terrain_env_golden.npz by running the REFERENCE's terrain trainer environment
(/root/reference/Code/mujoco/sim2real/train2.py:159-411 `QuadrupedEnv`, imported unmodified) with the third-party
`mujoco` module stubbed onto the oracle physics. Needs /root/reference (absent on the GPU box).

Pinned: the 8-action mapping and clipping, the 12-float observation, the twelve reward terms, the termination rules,
reset-with-settle, the terrain generator (`_generate_random_terrain`, module `random` seeded through the reference's own
`reset()` with a frozen clock) and `get_terrain_height`. Not pinned: the physics (oracle/odg_oracle.h).
The stub model mirrors what MuJoCo would compile from `walking_scene.xml`, the scene train2.py loads (:66,454): an
hfield ASSET 100 x 100 of size (5, 5, 0.3, 0.001) and NO hfield geom, no "obstacle" geom — the floor is the plane.
The script also asserts that oracle/sim2real_oracle.py reproduces every recorded value.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/mujoco"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from make_golden_sim2real import install_mujoco_stub  # noqa: E402
from oracle.sim2real_oracle import QuadrupedEnvV2Oracle, terrain_height  # noqa: E402


def extend_stub_for_train2():
    import mujoco
    obj = mujoco.mjtObj
    obj.mjOBJ_HFIELD, obj.mjOBJ_GEOM = 4, 5
    base_name2id = mujoco.mj_name2id

    def mj_name2id(model, objtype, name):
        if objtype == obj.mjOBJ_HFIELD:
            return 0 if name == "terrain_hfield" else -1           # walking_scene.xml:19
        if objtype == obj.mjOBJ_GEOM:
            return -1                                              # neither "terrain_hfield" nor "obstacle" geoms exist there
        return base_name2id(model, objtype, name)
    mujoco.mj_name2id = mj_name2id
    Model = mujoco.MjModel
    orig_init = Model.__init__

    def init(self):
        orig_init(self)
        self.hfield_nrow = np.array([100]); self.hfield_ncol = np.array([100])
        self.hfield_size = np.array([[5.0, 5.0, 0.3, 0.001]])
        self.hfield_data = np.zeros(100 * 100, dtype=np.float32)
        self.geom_pos = np.zeros((1, 3)); self.geom_rgba = np.ones((1, 4))
    Model.__init__ = init
    mujoco.mj_resetData = lambda model, data: data._sim.reset_keyframe() or data._sim.qpos.__setitem__(slice(None), data._sim.qpos * 0) or None
    Data = mujoco.MjData
    orig_dinit = Data.__init__

    def dinit(self, model):
        orig_dinit(self, model)
    Data.__init__ = dinit


def main():
    install_mujoco_stub()
    extend_stub_for_train2()
    sys.path.insert(0, os.path.join(REF, "sim2real"))
    import train2 as ref                                             # the reference module, unmodified
    import mujoco
    # the global maps train() fills in (train2.py:457-468), replayed verbatim on the stub model
    vm = mujoco.MjModel.from_xml_path("our_robot/walking_scene.xml"); vd = mujoco.MjData(vm)
    for an in ref.ACTUATOR_NAMES_ORDERED:
        aid = mujoco.mj_name2id(vm, mujoco.mjtObj.mjOBJ_ACTUATOR, an)
        jid = vm.actuator_trnid[aid, 0]; jn = mujoco.mj_id2name(vm, mujoco.mjtObj.mjOBJ_JOINT, jid)
        ref.ACTUATOR_TO_JOINT_NAME_MAP[an] = jn; ref.JOINT_NAME_TO_QPOS_IDX_MAP[jn] = vm.jnt_qposadr[jid]
    mujoco.mj_resetDataKeyframe(vm, vd, 0)
    for an in ref.ACTUATOR_NAMES_ORDERED:
        jn = ref.ACTUATOR_TO_JOINT_NAME_MAP[an]
        ref.sim_keyframe_home_qpos_map[jn] = vd.qpos[ref.JOINT_NAME_TO_QPOS_IDX_MAP[jn]]

    env = ref.QuadrupedEnv("our_robot/walking_scene.xml")
    assert env.hfield_id == 0 and env.obstacle_geom_id == -1 and env.sim_steps_per_policy_step == 40
    assert (env.state_dim, env.action_dim) == (12, 8)
    orc = QuadrupedEnvV2Oracle()
    rng = np.random.default_rng(11)
    n_ep, n_steps = 4, 30
    clocks = [1000.123, 1004.25, 1002.789, 1005.75]                  # frozen `time.time()` per reset -> the terrain seed (:341-343)
    rec = {k: [] for k in ("action", "obs", "reward", "done", "reason", "sim_target_rad", "reset_obs", "inject", "hfield", "seed")}
    reasons = ["max_steps", "mj_error", "orientation_limit", "too_much_backward"]
    for ep in range(n_ep):
        ref.time.time = lambda c=clocks[ep]: c
        seed = int(clocks[ep] * 1000) % (2 ** 32 - 1)
        ro = env.reset(); oo = orc.reset(seed=seed)
        assert np.array_equal(ro, oo), (ro, oo)
        assert np.array_equal(env.model.hfield_data, orc.hfield_data), ep
        per = {k: [] for k in rec}
        for t in range(n_steps):
            a = rng.uniform(-1, 1, 8).astype(np.float32)
            if ep == 1:
                a[[0, 2, 4, 6]] = -abs(a[0]) * np.array([1, 1, -1, -1])      # drive it backwards
            inj = np.full(4, np.nan)
            if ep == 2 and t == 3:                                   # roll the trunk 69 degrees: "orientation_limit"
                inj[:] = [np.cos(0.6), np.sin(0.6), 0.0, 0.0]
                env.data.qpos[3:7] = inj; orc.sim.qpos[3:7] = inj
            robs, rr, rd, rinfo = env.step(a.astype(np.float64))     # (float64: the NumPy 1.26 promotion the reference ran under)
            oobs, orr, od, oinfo = orc.step(a)
            assert np.array_equal(robs, oobs), (ep, t)
            assert abs(rr - orr) <= 1e-12 * max(1.0, abs(rr)), (ep, t, rr, orr)
            assert rd == od and rinfo["termination_reason"] == oinfo["termination_reason"], (ep, t)
            assert np.array_equal(rinfo["sim_target_rad"], oinfo["sim_target_rad"])
            per["action"].append(a); per["obs"].append(robs); per["reward"].append(rr); per["done"].append(rd)
            per["reason"].append(reasons.index(rinfo["termination_reason"])); per["sim_target_rad"].append(rinfo["sim_target_rad"])
            per["inject"].append(inj)
            if rd:
                break
        n = len(per["action"]); pad = n_steps - n
        for k in ("action", "obs", "reward", "done", "reason", "sim_target_rad", "inject"):
            arr = np.array(per[k])
            if pad:
                arr = np.concatenate([arr, np.zeros((pad,) + arr.shape[1:], arr.dtype)])
            rec[k].append(arr)
        rec["reset_obs"].append(ro); rec["hfield"].append(env.model.hfield_data.copy()); rec["seed"].append(seed)
        rec.setdefault("length", []).append(n)
    # get_terrain_height (:295-304) on the last generated non-flat terrain
    k = max(range(n_ep), key=lambda e: float(np.ptp(rec["hfield"][e])))
    env.model.hfield_data[:] = rec["hfield"][k]
    pts = rng.uniform(-2.6, 2.6, (64, 2))
    hts = np.array([env.get_terrain_height(x, y) for x, y in pts])
    assert np.array_equal(hts, np.array([terrain_height(rec["hfield"][k], x, y) for x, y in pts]))
    out = {key: (np.stack(v) if key not in ("length", "seed") else np.array(v)) for key, v in rec.items()}
    out["height_points"], out["heights"], out["height_terrain"] = pts, hts, np.array(k)
    path = os.path.join(ROOT, "tests", "golden", "terrain_env_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: lengths {out['length']}, final reasons {[int(out['reason'][e, out['length'][e] - 1]) for e in range(n_ep)]}, "
          f"terrain ranges {[float(np.ptp(h)) for h in out['hfield']]}")


if __name__ == "__main__":
    main()
