"""CPU restatement of the Go1 task environments' logic. TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

`JumpEnv` restates `JumpEnvironmentV0` (reference: Code/mujoco/environments/JumpEnvironment.py:70-134) and
`JumpEnvironmentRewardCalc` (rewards/jump_environment_reward_calc.py:55-182) on top of the oracle's physics
(`oracle.Sim("go1")`: the flat-floor Go1 scene — the reference's own jump_scene.xml nests the floor in a second body
named `trunk`, which MuJoCo rejects, so the scene the class was written for cannot be loaded by the reference either).
Pinned to the reference's own code by tests/golden/jump_env_golden.npz (tools/make_golden_jump.py imports the two
modules unmodified and runs them on the same physics); tests/test_golden_jump.py holds this file to it bit-exactly.

`landing_is_healthy` / `landing_obs` restate the two pieces of `LandingEnvironmentV0` that exist
(landing_environment_reward_calc.py:59-75, landing_environment.py:116-136); the class itself cannot be imported
(landing_environment.py:5) and its reward calls methods its calculator does not define (:90-109).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .oracle import Sim, lib, philox

K_STREAM_RESET, K_STREAM_DESVEL = 0x52534554, 0x44564C00
JUMP_TERMS = ("landing_precision", "landing_orientation", "control_velocity_horizontal", "height_clearance", "phase_sync",
              "jump_velocity", "distance_on_liftoff", "vertical_velocity_on_landing", "out_of_bounds", "collision_cost")
CONTACT_BODIES = [2, 3, 5, 6, 8, 9, 11, 12]          # jump_environment_reward_calc.py:49 (hips and thighs)
FEET_BODIES = [4, 7, 10, 13]                         # :48 (the calf bodies, which carry the foot spheres)


def _u01(x):
    return np.float32(x >> 8) * np.float32(5.9604644775390625e-08)


def euler_from_quaternion(q):
    """reward_calc euler_from_quaternion (jump :163-182), on the raw quaternion."""
    w, x, y, z = q
    roll = np.arctan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    t2 = 2.0 * (w * y - z * x)
    t2 = 1.0 if t2 > 1.0 else t2
    t2 = -1.0 if t2 < -1.0 else t2
    pitch = np.arcsin(t2)
    yaw = np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return roll, pitch, yaw


def projected_gravity(gravity, quat):
    e = np.array(euler_from_quaternion(quat))
    v = np.dot(gravity, e) * e
    n = np.linalg.norm(v)
    return v if n == 0 else v / n


def jump_reward_terms(qpos, qvel, cfrc_ext, desired_velocity):
    """compute_rewards (jump_environment_reward_calc.py:55-92): the ten weighted terms in reward_info's key order."""
    cube_xy, cube_h = np.array([1.0, 0.0]), 0.5
    dist = np.linalg.norm(cube_xy - qpos[:2])
    roll, pitch, yaw = euler_from_quaternion(qpos[3:7])
    t = np.zeros(10)
    t[0] = (np.exp(-dist) if qpos[2] >= cube_h else 0) * 3.0
    t[1] = np.exp(-(abs(roll) + abs(pitch) + abs(yaw))) * 2.0
    t[2] = np.exp(-np.linalg.norm(qvel[:2])) * 1.0
    t[3] = max(0, qpos[2] - cube_h) * .2
    t[4] = -0.0 * 0.8                                     # feet_air_time stays zeros(4): nothing advances it
    t[5] = np.exp(-np.sum(np.square(desired_velocity - qvel[:3])) / 0.45) * 1.0
    t[6] = (np.exp(dist) if qpos[2] < cube_h else 0) * 2.0
    t[7] = (qvel[2] ** 2 if qpos[2] >= cube_h else 0) * 1.5
    t[8] = (1.0 if dist > 1.0 else 0) * 3.0
    t[9] = np.sum(1.0 * (np.linalg.norm(cfrc_ext[CONTACT_BODIES]) > 0.1)) * 1.0
    return t


def jump_reward(terms):
    rewards = terms[0] + terms[1] + terms[2] + terms[3] + terms[4] + terms[5]
    costs = terms[6] + terms[7] + terms[8] + terms[9]
    return max(0.0, rewards - costs), rewards - costs


def static_stability(qpos, qvel):
    """:140-150 — finite state, roll and yaw inside +-20 degrees (inclusive)."""
    state = np.concatenate([qpos, qvel])
    roll, _, yaw = euler_from_quaternion(state[3:7])
    lim = np.deg2rad(20)
    return bool(np.isfinite(state).all() and -lim <= yaw <= lim and -lim <= roll <= lim)


class JumpEnv:
    def __init__(self, seed=0, env_id=0, model="go1", max_steps=750, frame_skip=10, noise=0.1):
        self.sim = Sim(model)
        self.seed, self.env_id, self.episode = seed, env_id, 0
        self.max_steps, self.frame_skip, self.noise = max_steps, frame_skip, np.float32(noise)
        self.step_count = 0
        r = philox(seed, env_id, 0, 0, K_STREAM_DESVEL)
        f = np.float32
        self.desired_velocity = np.array([f(1.20) + f(0.05) * _u01(r[0]), 0.0, f(1.20) + f(0.05) * _u01(r[1])], np.float32)
        self.gravity = np.array(self.sim.desc["gravity"])

    @property
    def qpos(self): return self.sim.qpos
    @property
    def qvel(self): return self.sim.qvel

    def obs(self):
        q, v = self.sim.qpos, self.sim.qvel
        o = np.concatenate(([0.3 - q[0]], [0.3 - q[2]], v[:3], [v[2]], projected_gravity(self.gravity, q[3:7]), np.zeros(12)))
        return o.clip(-100.0, 100.0)

    def reset(self):
        """mj_resetData + reset_model (JumpEnvironment.py:119-134) with the counter-based noise of the CUDA library."""
        lib().odgo_reset_data(C.byref(self.sim.m), C.byref(self.sim.d))
        nq = self.sim.nq
        key = np.array(self.sim.desc["key_qpos"], np.float32)
        out = np.zeros(nq, np.float32)
        for blk in range((nq + 3) // 4):
            r = philox(self.seed, self.env_id, self.episode, blk, K_STREAM_RESET)
            for k in range(4):
                i = blk * 4 + k
                if i < nq:
                    nz = np.float32(-self.noise) + np.float32(np.float32(2.0) * self.noise) * _u01(r[k])
                    out[i] = key[i] + np.float32(nz)
        self.sim.qpos[:] = out
        self.sim.ctrl[:] = self.sim.desc["key_ctrl"]
        self.episode += 1
        self.step_count = 0
        return self.obs()

    def evaluate(self, cfrc=None):
        """Outputs on the current state (after one forward pass when cfrc is None)."""
        if cfrc is None:
            self.sim.forward()
            cfrc = self.sim.cfrc_ext()
        q, v = self.sim.qpos.copy(), self.sim.qvel.copy()
        terms = jump_reward_terms(q, v, cfrc, self.desired_velocity.astype(np.float64))
        reward, raw = jump_reward(terms)
        info = dict(x_position=q[0], y_position=q[1], z_position=q[1], distance_from_origin=np.linalg.norm(q[0:2], ord=2),
                    terms=terms, reward_unclipped=raw, cfrc_ext=cfrc,
                    collision_norm=float(np.linalg.norm(cfrc[CONTACT_BODIES])))
        return self.obs(), reward, (not static_stability(q, v)), self.step_count >= self.max_steps, info

    def step(self, ctrl):
        self.step_count += 1
        self.sim.ctrl[:] = ctrl
        for _ in range(self.frame_skip):
            self.sim.step()
        return self.evaluate(self.sim.cfrc_ext())


def landing_is_healthy(qpos, qvel):
    """landing_environment_reward_calc.py:59-75: finite, z in [0.22, 0.65], roll / pitch / yaw inside +-10 degrees."""
    state = np.concatenate([qpos, qvel])
    roll, pitch, yaw = euler_from_quaternion(state[3:7])
    lim = np.deg2rad(10)
    ok = np.isfinite(state).all() and 0.22 <= state[2] <= 0.65
    return bool(ok and -lim <= yaw <= lim and -lim <= roll <= lim and -lim <= pitch <= lim)
