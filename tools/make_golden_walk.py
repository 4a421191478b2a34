#!/usr/bin/env python
"""Generate tests/golden/walk_env_golden.npz by running the REFERENCE's own Python env code.

    python tools/make_golden_walk.py            # needs /root/reference (not present on the GPU box)

What is pinned: everything the reference itself implements for the walk path —
`ScaleActionWrapper.action` (environments/ScaleActionEnvironment.py:21-23),
`WalkEnvironmentV0.step/_get_obs/_calculate_rewards/reset_model` (environments/WalkEnvironment.py:56-151)
and `WalkEnvironmentRewardCalc` (rewards/walk_environment_reward_calc.py), including its stateful
quirks (double `diagonal_gait_reward` call, never-reset gait state, broadcast joint offset, ...).
These modules are imported UNMODIFIED from /root/reference and executed here.

What is NOT pinned: the physics. `mujoco`, `gymnasium` are third-party wheels that are absent, so they
are replaced by stubs: `MujocoEnv.do_simulation` advances oracle/odg_oracle.c (our fp64 restatement of
mj_step) and `mujoco.mj_contactForce` / `mju_quat2Mat` read the oracle's data. The golden file
therefore fixes "reference env logic on top of a given physics state sequence".

Two adaptations, both recorded in the file's `notes`:
 * NumPy: the reference ran NumPy 1.26 (best_model.zip system_info), where `np.float32 * python_float`
   gives float64; this container has NumPy >= 2 (NEP 50: float32). The reward/cost weight dicts are
   re-wrapped as np.float64 so the arithmetic is the 1.26 one.
 * RNG: `reset_model` noise and the (unseeded) desired velocity come from numpy generators that cannot be
   reproduced on the GPU; the script overwrites them with the Philox-derived values the oracle/GPU use.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/mujoco"
sys.path.insert(0, ROOT)

from oracle.oracle import Sim, WalkEnv  # noqa: E402
from opendog_b200.model.mjcf import quat_mul  # noqa: E402


# ----------------------------------------------------------------------------- stubs of the 3P modules
def install_stubs():
    mj = types.ModuleType("mujoco")

    def mju_quat2Mat(res, quat):
        w, x, y, z = quat
        res[:] = [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                  2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                  2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]
    mj.mju_quat2Mat = mju_quat2Mat
    fn = types.ModuleType("mujoco._functions")

    def mj_contactForce(model, data, idx, out):
        out[:] = 0
        out[0:3] = data._sim.contacts()[idx]["force"]
    fn.mj_contactForce = mj_contactForce
    mj._functions = fn
    sys.modules["mujoco"] = mj
    sys.modules["mujoco._functions"] = fn

    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype
    spaces.Box = Box
    gym.spaces = spaces

    class ActionWrapper:
        def __init__(self, env):
            self.env = env
            self.unwrapped = getattr(env, "unwrapped", env)

        def step(self, action):
            return self.env.step(self.action(action))

        def reset(self, **kw):
            return self.env.reset(**kw)
    gym.ActionWrapper = ActionWrapper
    envs = types.ModuleType("gymnasium.envs")
    gmj = types.ModuleType("gymnasium.envs.mujoco")

    class _Contact:
        def __init__(self, g1, g2, frame):
            self.geom1, self.geom2, self.frame = g1, g2, frame

    class _Data:
        def __init__(self, sim):
            self._sim = sim
            self.qpos, self.qvel, self.ctrl = sim.qpos, sim.qvel, sim.ctrl

        @property
        def ncon(self):
            return self._sim.ncon

        @property
        def contact(self):
            geoms = self._sim.desc["geoms"]
            out = []
            for c in range(self._sim.d.ncon):
                k = self._sim.d.contact[c]
                out.append(_Contact(0, geoms[k.geom]["mj_geom_id"], np.array(k.frame[:])))
            return out

        @property
        def qfrc_actuator(self):
            return self._sim.qfrc_actuator

        @property
        def time(self):
            return self._sim.d.time

        @property
        def xquat(self):
            """[nbody,4] in MuJoCo body order: world, trunk, then per leg (thigh, calf, paw)."""
            d = self._sim.d
            xq = np.array(d.xquat[:]).reshape(-1, 4)
            paw_quat = np.array([0.0, -0.38268343, 0.0, 0.92387953])
            paw_quat = paw_quat / np.linalg.norm(paw_quat)
            rows = [np.array([1.0, 0, 0, 0]), xq[0]]
            for leg in range(4):
                rows += [xq[1 + 2 * leg], xq[2 + 2 * leg], quat_mul(xq[2 + 2 * leg], paw_quat)]
            return np.stack(rows)

    class _Opt:
        pass

    class _Model:
        def __init__(self, desc):
            self.opt = _Opt()
            self.opt.gravity = np.array(desc["gravity"])
            self.opt.timestep = desc["timestep"]
            self.key_ctrl = np.array([desc["key_ctrl"]])
            self.key_qpos = np.array([desc["key_qpos"]])
            self.actuator_ctrlrange = np.array(desc["act_ctrlrange"])
            self.nq, self.nv = desc["nq"], desc["nv"]
            gb = np.zeros(14, dtype=int)
            gb[1] = 1
            for g in desc["geoms"]:
                gb[g["mj_geom_id"]] = g["mj_body_id"]
            self.geom_bodyid = gb

    class MujocoEnv:
        def __init__(self, model_path, frame_skip, observation_space, default_camera_config=None, **kw):
            assert model_path.endswith("our_robot/walking_scene.xml")
            self._sim = Sim("our_robot")
            self.model = _Model(self._sim.desc)
            self.data = _Data(self._sim)
            self.frame_skip = frame_skip
            self.np_random = np.random.default_rng(0)
            self.unwrapped = self

        @property
        def dt(self):
            return self.model.opt.timestep * self.frame_skip

        def do_simulation(self, ctrl, n_frames):
            self.data.ctrl[:] = ctrl
            for _ in range(n_frames):
                self._sim.step()

        def reset(self, seed=None, options=None):
            import ctypes as C
            from oracle.oracle import lib
            lib().odgo_reset_data(C.byref(self._sim.m), C.byref(self._sim.d))     # mj_resetData
            ob = self.reset_model()
            return ob, self._get_reset_info()

        def render(self):
            pass
    gmj.MujocoEnv = MujocoEnv
    envs.mujoco = gmj
    gym.envs = envs
    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", spaces), ("gymnasium.envs", envs),
                      ("gymnasium.envs.mujoco", gmj)):
        sys.modules[name] = mod


def main():
    install_stubs()
    sys.path.insert(0, REF)
    from environments.WalkEnvironment import WalkEnvironmentV0          # the reference, unmodified
    from environments.ScaleActionEnvironment import ScaleActionWrapper

    n_envs, n_steps, max_steps, seed = 4, 90, 30, 123
    rng = np.random.default_rng(2024)
    rec = {k: [] for k in ("action", "obs", "reward", "terminated", "truncated", "x_position", "y_position",
                           "distance_from_origin", "patterns_matches", "paw_contact_forces",
                           "linear_vel_tracking_reward", "reward_ctrl", "reset_obs", "did_reset", "inject", "gait_first")}
    init_qpos, init_desvel = [], []
    max_dev = 0.0
    for i in range(n_envs):
        ref_core = WalkEnvironmentV0()
        ref_core._max_episode_time_sec = max_steps * ref_core.dt          # shorten episodes to exercise truncation
        ref = ScaleActionWrapper(ref_core)
        # NumPy-1.26 scalar promotion (see module docstring)
        ref_core.utils.cost_weights = {k: np.float64(v) for k, v in ref_core.utils.cost_weights.items()}
        ref_core.utils.reward_weights = {k: np.float64(v) for k, v in ref_core.utils.reward_weights.items()}
        w = WalkEnv(seed=seed, env_id=i)
        w.e.max_steps = max_steps
        ref_core.utils.desired_velocity = np.array(w.desired_velocity)
        init_desvel.append(np.array(w.desired_velocity))

        def sync_reset():
            oobs = w.reset()
            ref.reset()
            ref_core.data.qpos[:] = w.qpos                 # inject the Philox reset state
            robs = ref_core._get_obs()
            assert np.array_equal(robs, oobs)
            return robs
        first = sync_reset()
        init_qpos.append(w.qpos.copy())
        per = {k: [] for k in rec}
        per["reset_obs"].append(first)
        for t in range(n_steps):
            a = rng.uniform(-1, 1, 8).astype(np.float32)
            if t % 11 == 0:
                a = np.clip(a * 3, -1, 1).astype(np.float32)   # saturated actions too
            # state injections (applied to BOTH sides before the step) so that termination and the
            # gait-pattern streak logic are exercised; recorded so the test can replay them
            inj = np.full(5, np.nan)
            if i == 1 and t in (20, 47):
                ang = 0.33 if t == 20 else -0.4
                inj[:4] = [np.cos(ang / 2), np.sin(ang / 2) if t == 20 else 0.0, 0.0 if t == 20 else np.sin(ang / 2), 0.0]
            if i == 2 and 38 <= t < 70:
                inj[4] = 2.5
            if i == 3 and t % 9 == 4:
                inj[4] = 1.8
            for side_q, side_v in ((ref_core.data.qpos, ref_core.data.qvel), (w.qpos, w.qvel)):
                if not np.isnan(inj[0]):
                    side_q[3:7] = inj[:4]
                if not np.isnan(inj[4]):
                    side_v[0] = inj[4]
            per["inject"].append(inj)
            robs, rrew, rterm, rtrunc, rinfo = ref.step(a)
            oobs, orew, oterm, otrunc, oinfo = w.step(a)
            # the oracle must reproduce the reference's outputs on the same physics
            assert np.array_equal(robs, oobs), (i, t)
            assert bool(rterm) == oterm and bool(rtrunc) == otrunc, (i, t)
            assert rinfo["patterns_matches"] == oinfo["patterns_matches"]
            max_dev = max(max_dev, abs(rrew - orew))
            pf = np.stack([rinfo["paw_contact_forces"][b] for b in (4, 7, 10, 13)])
            assert np.allclose(pf, oinfo["paw_contact_forces"], rtol=0, atol=1e-12)
            per["action"].append(a); per["obs"].append(robs); per["reward"].append(rrew)
            per["terminated"].append(bool(rterm)); per["truncated"].append(bool(rtrunc))
            for k in ("x_position", "y_position", "distance_from_origin", "patterns_matches",
                      "linear_vel_tracking_reward", "reward_ctrl"):
                per[k].append(float(rinfo[k]))
            per["paw_contact_forces"].append(pf)
            per["gait_first"].append(oinfo["gait_first_call"])
            done = bool(rterm) or bool(rtrunc)
            per["did_reset"].append(done)
            if done:
                per["reset_obs"].append(sync_reset())
            else:
                per["reset_obs"].append(np.zeros(33))
        per["reset_obs"] = per["reset_obs"][1:]
        for k in rec:
            rec[k].append(np.array(per[k]))
    out = {k: np.stack(v) for k, v in rec.items()}
    out["init_qpos"] = np.stack(init_qpos); out["desired_velocity"] = np.stack(init_desvel)
    out["seed"] = seed; out["max_steps"] = max_steps
    out["notes"] = np.array("reference env code (WalkEnvironment.py, ScaleActionEnvironment.py, "
                            "walk_environment_reward_calc.py) on oracle physics; numpy-1.26 promotion; Philox reset")
    path = os.path.join(ROOT, "tests", "golden", "walk_env_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {n_envs} envs x {n_steps} steps, resets={int(out['did_reset'].sum())}, "
          f"terminated={int(out['terminated'].sum())}, gait matches={int((out['gait_first'] > 0).sum())}, "
          f"max streak={int(out['gait_first'].max())}, max |reward ref - oracle| = {max_dev:.3e}")


if __name__ == "__main__":
    main()
