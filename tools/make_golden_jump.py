#!/usr/bin/env python
"""Generate tests/golden/jump_env_golden.npz by running the REFERENCE's own Jump environment code.

    python tools/make_golden_jump.py            # needs /root/reference (not present on the GPU box)

Pinned: `JumpEnvironmentV0.step/_get_obs` (environments/JumpEnvironment.py:70-117) and
`JumpEnvironmentRewardCalc` (rewards/jump_environment_reward_calc.py), imported UNMODIFIED from /root/reference,
including their quirks (obs reads `utils.last_action`, which nothing updates; `feet_air_time` never advances;
info["z_position"] is qpos[1]; the collision cost thresholds the Frobenius norm of eight cfrc_ext rows at once).

Not pinned: the physics. `mujoco` / `gymnasium` are absent third-party wheels and are replaced by stubs whose
`MujocoEnv.do_simulation` advances oracle/odg_oracle.c on the flat-floor Go1 scene and whose `data.cfrc_ext` is the
oracle's mj_rnePostConstraint restatement. (The reference's own jump_scene.xml declares a second body named `trunk`
with the floor nested inside it — MuJoCo refuses to compile it — so no other scene exists for this class.)
RNG: reset noise and the unseeded desired velocity are overwritten with the Philox-derived values the oracle / GPU use.
NumPy >= 2 vs the reference's 1.26: every operand here is already float64, so promotion rules do not differ.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/mujoco"
sys.path.insert(0, ROOT)

from oracle.oracle import Sim  # noqa: E402
from oracle.go1_tasks import JUMP_TERMS, JumpEnv  # noqa: E402


def install_stubs():
    mj = types.ModuleType("mujoco")
    mj.mj_name2id = lambda model, kind, name: 0

    class _V:
        value = 0

    class mjtObj:
        mjOBJ_SITE = _V()
        mjOBJ_BODY = _V()
    mj.mjtObj = mjtObj
    sys.modules["mujoco"] = mj

    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype
    spaces.Box = Box
    gym.spaces = spaces
    envs = types.ModuleType("gymnasium.envs")
    gmj = types.ModuleType("gymnasium.envs.mujoco")

    class _Data:
        def __init__(self, sim):
            self._sim = sim
            self.qpos, self.qvel, self.ctrl = sim.qpos, sim.qvel, sim.ctrl
            self.cfrc_ext = np.zeros((14, 6))

        @property
        def time(self):
            return self._sim.d.time

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

    class MujocoEnv:
        def __init__(self, model_path, frame_skip, observation_space, default_camera_config=None, **kw):
            assert model_path.endswith("unitree_go1/jump_scene.xml")
            self._sim = Sim("go1")
            self.model = _Model(self._sim.desc)
            self.data = _Data(self._sim)
            self.frame_skip = frame_skip
            self.np_random = np.random.default_rng(0)
            self.render_mode = None

        @property
        def dt(self):
            return self.model.opt.timestep * self.frame_skip

        def do_simulation(self, ctrl, n_frames):
            self.data.ctrl[:] = ctrl
            for _ in range(n_frames):
                self._sim.step()
            self.data.cfrc_ext = self._sim.cfrc_ext()          # mj_rnePostConstraint

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
    from environments.JumpEnvironment import JumpEnvironmentV0          # the reference, unmodified

    n_envs, n_steps, max_steps, seed = 4, 60, 25, 77
    rng = np.random.default_rng(99)
    keys = ("action", "inject_qpos", "inject_qvel", "obs", "reward", "terminated", "truncated", "x_position", "y_position",
            "z_position", "distance_from_origin", "terms", "did_reset", "reset_obs", "qpos", "qvel", "collision_norm")
    rec = {k: [] for k in keys}
    desvel = []
    for i in range(n_envs):
        ref = JumpEnvironmentV0(None)
        ref._max_episode_time_sec = max_steps * ref.dt
        w = JumpEnv(seed=seed, env_id=i, max_steps=max_steps)
        ref.utils.desired_velocity = w.desired_velocity.astype(np.float64)
        desvel.append(w.desired_velocity.copy())
        lo, hi = ref.model.actuator_ctrlrange[:, 0], ref.model.actuator_ctrlrange[:, 1]

        def sync_reset():
            oobs = w.reset()
            ref._sim.reset_keyframe()
            ref.data.qpos[:] = w.qpos                            # inject the Philox reset state
            ref.data.qvel[:] = 0
            ref._step = 0
            robs = ref._get_obs()
            assert np.array_equal(robs, oobs)
            return robs
        sync_reset()
        per = {k: [] for k in keys}
        for t in range(n_steps):
            a = np.clip(np.array(ref.model.key_ctrl[0]) + rng.uniform(-1, 1, 12) * (0.1 if i == 0 else 0.6), lo, hi)
            # state injections (both sides, before the step): airborne above the cube height, over the cube, tilted
            iq, iv = np.full(19, np.nan), np.full(18, np.nan)
            if i == 1 and t % 7 == 3:
                iq[:3] = [rng.uniform(0.2, 1.4), rng.uniform(-0.4, 0.4), rng.uniform(0.48, 0.7)]
                iv[:3] = rng.uniform(-1, 1.5, 3)
            if i == 2 and t in (12, 31):
                ang = 0.8 if t == 12 else -0.75
                iq[3:7] = [np.cos(ang / 2), np.sin(ang / 2) if t == 12 else 0, 0, 0 if t == 12 else np.sin(ang / 2)]
            if i == 3 and t % 10 == 5:
                iq[2] = 0.13; iq[3:7] = [np.cos(0.725), np.sin(0.725), 0.0, 0.0]  # on its side: hips and thighs touch
            for side_q, side_v in ((ref.data.qpos, ref.data.qvel), (w.qpos, w.qvel)):
                side_q[~np.isnan(iq)] = iq[~np.isnan(iq)]
                side_v[~np.isnan(iv)] = iv[~np.isnan(iv)]
            robs, rrew, rterm, rtrunc, rinfo = ref.step(a)
            oobs, orew, oterm, otrunc, oinfo = w.step(a)
            assert np.array_equal(robs, oobs), (i, t)
            assert bool(rterm) == oterm and bool(rtrunc) == otrunc and rrew == orew, (i, t, rrew, orew)
            per["action"].append(a); per["inject_qpos"].append(iq); per["inject_qvel"].append(iv)
            per["obs"].append(robs); per["reward"].append(rrew)
            per["terminated"].append(bool(rterm)); per["truncated"].append(bool(rtrunc))
            for k in ("x_position", "y_position", "z_position", "distance_from_origin"):
                per[k].append(float(rinfo[k]))
            per["terms"].append([float(rinfo[k]) for k in JUMP_TERMS])
            per["qpos"].append(w.qpos.copy()); per["qvel"].append(w.qvel.copy())
            per["collision_norm"].append(oinfo["collision_norm"])
            done = bool(rterm) or bool(rtrunc)
            per["did_reset"].append(done)
            per["reset_obs"].append(sync_reset() if done else np.zeros(21))
        for k in keys:
            rec[k].append(np.array(per[k]))
    out = {k: np.stack(v) for k, v in rec.items()}
    out["desired_velocity"] = np.stack(desvel); out["seed"] = seed; out["max_steps"] = max_steps
    out["notes"] = np.array("reference JumpEnvironment.py + jump_environment_reward_calc.py on oracle physics (flat-floor Go1); "
                            "Philox reset noise / desired velocity")
    path = os.path.join(ROOT, "tests", "golden", "jump_env_golden.npz")
    np.savez_compressed(path, **out)
    t = out["terms"]
    print(f"wrote {path}: {n_envs} x {n_steps} steps, resets {int(out['did_reset'].sum())}, terminated {int(out['terminated'].sum())}, "
          f"reward>0 {int((out['reward'] > 0).sum())}, above cube {int((t[..., 0] > 0).sum())}, collision {int((t[..., 9] > 0).sum())}, "
          f"out of bounds {int((t[..., 8] > 0).sum())}")


if __name__ == "__main__":
    main()
