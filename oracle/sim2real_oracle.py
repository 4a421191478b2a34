"""CPU restatement of `QuadrupedEnv` (/root/reference/Code/mujoco/sim2real/train.py:151-411) on the oracle
physics. TEST INFRASTRUCTURE ONLY (tests/, smoke, bench CPU legs). fp64, one environment, plain Python.

Pinned by tests/golden/sim2real_env_golden.npz: tools/make_golden_sim2real.py runs the reference's own class
(imported unmodified, `mujoco` stubbed onto the oracle physics) on the same action sequences.
"""
from __future__ import annotations

import math

import numpy as np

from .oracle import Sim

ORDERED = ["FR_tigh_actuator", "FR_knee_actuator", "FL_tigh_actuator", "FL_knee_actuator",
           "BR_tigh_actuator", "BR_knee_actuator", "BL_tigh_actuator", "BL_knee_actuator"]      # train.py:25-30
REAL_HOME_DEG = [-45.0, 45.0, 45.0, 45.0, 45.0, -45.0, 45.0, -45.0]                            # train.py:95-102
AMP = math.radians(40.0)                                                                         # train.py:75-76


def quat_to_ypr(q):                                                                              # train.py:110-118
    q0, q1, q2, q3 = q
    roll = math.atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2))
    sinp = 2 * (q0 * q2 - q3 * q1)
    pitch = math.asin(sinp) if abs(sinp) < 1 else math.copysign(math.pi / 2, sinp)
    yaw = math.atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))
    return yaw, pitch, roll


class QuadrupedEnvOracle:
    def __init__(self, model="our_robot"):
        self.sim = Sim(model)
        d = self.sim.desc
        names = d["act_names"]
        self.act_id = [names.index(n) for n in ORDERED]
        # joint of each actuator: qpos index 7 + leg*njl + joint
        self.qidx = [7 + d["act_leg"][u] * d["njl"] + d["act_joint"][u] for u in self.act_id]
        self.vidx = [q - 1 for q in self.qidx]
        self.home = [d["key_qpos"][q] for q in self.qidx]
        self.ctrlrange = [d["act_ctrlrange"][u] for u in self.act_id]
        self.initial_ctrl = np.array(d["key_ctrl"], dtype=np.float64)
        self.initial_y = d["key_qpos"][1]
        self.n_sub = max(1, int(0.10 / d["timestep"]))
        self.state_dim, self.action_dim = 22, 4
        self.counter = 0
        self.last_cmd = self.initial_ctrl.copy()
        self.prev_x = 0.0; self.cpos = 0.0; self.cneg = 0.0; self.prev_net = 0.0

    def obs(self):                                                                               # train.py:184-207
        s = self.sim
        yaw, pitch, roll = quat_to_ypr(s.qpos[3:7])
        jp = [s.qpos[q] - h for q, h in zip(self.qidx, self.home)]
        jv = [s.qvel[v] for v in self.vidx]
        ph = (self.counter % 2) / 1.0
        return np.concatenate([[yaw, pitch, roll], jp, jv, [s.qvel[0]], [np.sin(ph * np.pi), np.cos(ph * np.pi)]]).astype(np.float32)

    def reset(self):                                                                             # train.py:209-233
        s = self.sim
        self.counter = 0
        s.reset_keyframe()
        s.ctrl[:] = self.initial_ctrl
        for _ in range(100):
            s.ctrl[:] = self.initial_ctrl
            s.step()
        self.last_cmd = s.ctrl.copy()
        self.prev_x = float(s.qpos[0]); self.cpos = self.cneg = self.prev_net = 0.0
        return self.obs()

    def step(self, a):                                                                           # train.py:287-411
        s = self.sim
        phase = self.counter % 2
        fr, k1, fl, k2 = [float(x) * AMP for x in a]
        d = [fr, 0.0, fl, 0.0, fl, 0.0, fr, 0.0]
        if phase == 0:
            d[1], d[7] = k1, -k1
        else:
            d[3], d[5] = k2, -k2
        cmd = np.zeros(8)
        for o in range(8):
            cmd[self.act_id[o]] = np.clip(self.home[o] + d[o], self.ctrlrange[o][0], self.ctrlrange[o][1])
        s.ctrl[:] = cmd
        self.min_gap = np.full(3, 1e30)            # closest call of any discrete decision in this policy step (Sim.decision_gaps)
        for _ in range(self.n_sub):
            s.step()
            self.min_gap = np.minimum(self.min_gap, s.decision_gaps)
        mj_err = not (np.isfinite(s.qpos).all() and np.isfinite(s.qvel).all())
        self.counter += 1
        x = float(s.qpos[0]); dx = x - self.prev_x
        if dx > 0:
            self.cpos += dx
        elif dx < 0:
            self.cneg += abs(dx)
        self.prev_x = x
        obs = self.obs()
        vx = float(s.qvel[0])
        r = 150.0 * vx
        net = self.cpos - self.cneg; dnet = net - self.prev_net; self.prev_net = net
        if dnet > 0.0005:
            r += 15.0 * dnet
        if vx < -0.005:
            r += -5.0 * abs(vx)
        r += 0.05 + -0.2 * abs(s.qvel[1]) + -0.1 * abs(s.qpos[1] - self.initial_y)
        yaw, pitch, roll = quat_to_ypr(s.qpos[3:7])
        th, thy, lim = math.radians(5.0), math.radians(10.0), math.radians(25.0)
        ro = 0.0
        if abs(roll) > th: ro += -0.05 * (abs(roll) - th) ** 2
        if abs(pitch) > th: ro += -0.05 * (abs(pitch) - th) ** 2
        if abs(yaw) > thy: ro += -0.05 * (abs(yaw) - thy) ** 2
        r += ro
        r += -0.01 * float(np.sum([(cmd[u] - self.last_cmd[u]) ** 2 for u in self.act_id]))
        too_far = not_home = 0
        for leg in range(4):                       # FR FL BR BL
            swinging = (leg in (0, 3)) if phase == 0 else (leg in (1, 2))
            devs = [abs(math.degrees(cmd[self.act_id[leg * 2 + j]] - self.home[leg * 2 + j])) for j in range(2)]
            if swinging:
                too_far += max(devs) > 40.0
            else:
                not_home += any(dv > 15.0 for dv in devs)
        if too_far or not_home:
            r += -(too_far + not_home) * 0.5
        done, reason = False, "max_steps"
        if mj_err:
            r -= 20.0; done = True; reason = "mj_error"
        if abs(roll) > lim or abs(pitch) > lim or abs(yaw) > lim:
            r -= 5.0; done = True; reason = "orientation_limit"
        if not done and self.cpos > 0.05 and self.cneg > 0.75 * self.cpos:
            r -= 5.0; done = True; reason = "too_much_backward"
        self.last_cmd = cmd.copy()
        return obs, float(r), done, {"sim_target_rad": cmd.copy(), "termination_reason": reason}


# ================================================================================================================
# The terrain trainer's environment: `QuadrupedEnv` of /root/reference/Code/mujoco/sim2real/train2.py:159-411.
# As shipped it loads `walking_scene.xml` (train2.py:66,454): the height field there is an ASSET without a geom
# (walking_scene.xml:19,25 — the floor is the plane), so the terrain generated at every reset (:203-293) is written to
# `model.hfield_data`, can be queried (`get_terrain_height`, :295-304) and rendered, but never touches the physics;
# `walking_scene_terrain.xml`, which does have an hfield geom, is loaded by no script. What changes against train.py is
# the task: 8 direct joint targets (amplitude 50 deg), 12-float observation, 40 substeps per policy step, a 12-term
# reward, wider termination limits. Pinned by tests/golden/terrain_env_golden.npz (tools/make_golden_terrain.py).
import random as _random

AMP2 = math.radians(50.0)                                                                        # train2.py:93-94
TERRAIN_ROWS = TERRAIN_COLS = 100                                                                # :111-112
TERRAIN_MAX_ABS_HEIGHT, TERRAIN_SMOOTHNESS_FACTOR, TERRAIN_NUM_SMOOTH_PASSES = 1.5, 0.3, 4       # :113-115
HFIELD_SIZE = (5.0, 5.0, 0.3, 0.001)                                                             # walking_scene.xml:19
HFIELD_GEOM_POS = (0.0, 0.0, 0.0)                                                                # no hfield geom: train2.py:173


def generate_terrain(rng=_random, robot_start=(0.0, 0.0), hfield_pos=HFIELD_GEOM_POS, size=HFIELD_SIZE):
    """`_generate_random_terrain` (train2.py:203-293) with the module `random` replaced by `rng` (same draw order):
    returns hfield_data, float32 [ncol * nrow] (the normalised heights, TRANSPOSED before flattening, :270)."""
    R, C = TERRAIN_ROWS, TERRAIN_COLS
    if rng.random() < 0.5:                                                                       # :206-210
        return np.full(R * C, 0.5, dtype=np.float32)
    raw = np.zeros((R, C), dtype=np.float32)
    x_ext, y_ext = size[0], size[1]
    csx, csy = x_ext / (C - 1), y_ext / (R - 1)
    sx, sy = robot_start
    radius = rng.uniform(0.1, 0.4)
    H = TERRAIN_MAX_ABS_HEIGHT
    for r in range(R):
        for c in range(C):
            wx = hfield_pos[0] - (x_ext / 2.0) + c * csx
            wy = hfield_pos[1] - (y_ext / 2.0) + r * csy
            dist = np.sqrt((wx - sx) ** 2 + (wy - sy) ** 2)
            if dist >= radius:
                base = rng.uniform(-H, H)
                fx = rng.uniform(0.2, 0.6); fy = rng.uniform(0.2, 0.6)
                noise = (np.sin(wx * fx) * np.cos(wy * fy) + np.sin(wx * fx * 2) * np.cos(wy * fy * 2)) * H * 0.7
                spike = 0
                if rng.random() < 0.2:
                    spike = rng.uniform(-H * 0.8, H * 0.8)
                raw[r, c] = base + noise + spike
                if abs(dist - radius) < 1.0:
                    raw[r, c] *= 1.5
    return smooth_and_normalise(raw, radius, robot_start, hfield_pos, size)


def smooth_and_normalise(raw, radius, robot_start=(0.0, 0.0), hfield_pos=HFIELD_GEOM_POS, size=HFIELD_SIZE):
    """The deterministic tail of `_generate_random_terrain` (train2.py:250-270): 4 passes of a 3x3 box blend outside the
    flat circle (interior cells only), min-max normalisation, transpose, flatten."""
    R, C = raw.shape
    csx, csy = size[0] / (C - 1), size[1] / (R - 1)
    sm = raw.copy()
    for _ in range(TERRAIN_NUM_SMOOTH_PASSES):
        tmp = sm.copy()
        for r in range(1, R - 1):
            for c in range(1, C - 1):
                wx = hfield_pos[0] - (size[0] / 2.0) + c * csx
                wy = hfield_pos[1] - (size[1] / 2.0) + r * csy
                dist = np.sqrt((wx - robot_start[0]) ** 2 + (wy - robot_start[1]) ** 2)
                if dist >= radius:
                    avg = np.mean(tmp[r - 1:r + 2, c - 1:c + 2])
                    sm[r, c] = tmp[r, c] * (1 - TERRAIN_SMOOTHNESS_FACTOR) + avg * TERRAIN_SMOOTHNESS_FACTOR
    lo, hi = np.min(sm), np.max(sm)
    norm = np.full_like(sm, 0.5) if hi <= lo + 1e-4 else (sm - lo) / (hi - lo)
    return norm.T.flatten().astype(np.float32)


def terrain_height(hfield_data, wx, wy, hfield_pos=HFIELD_GEOM_POS, size=HFIELD_SIZE):
    """`get_terrain_height` (train2.py:295-304), including its `c * nrow + r` index into the transposed data."""
    C, R = TERRAIN_COLS, TERRAIN_ROWS
    lx, ly = wx - hfield_pos[0], wy - hfield_pos[1]
    cf = (lx + size[0] / 2.0) / size[0] * (C - 1); rf = (ly + size[1] / 2.0) / size[1] * (R - 1)
    c = int(np.clip(cf, 0, C - 1)); r = int(np.clip(rf, 0, R - 1))
    h_off = size[3] + float(hfield_data[c * R + r]) * size[2]        # float32 sample widened to float64, as numpy does there
    return hfield_pos[2] + h_off


class QuadrupedEnvV2Oracle:
    """train2.py `QuadrupedEnv` on the oracle physics (flat plane: see the note above)."""

    def __init__(self, model="our_robot", terrain=True):
        self.sim = Sim(model)
        d = self.sim.desc
        names = d["act_names"]
        self.act_id = [names.index(n) for n in ORDERED]
        self.qidx = [7 + d["act_leg"][u] * d["njl"] + d["act_joint"][u] for u in self.act_id]
        self.home = [d["key_qpos"][q] for q in self.qidx]
        self.ctrlrange = [d["act_ctrlrange"][u] for u in self.act_id]
        self.initial_ctrl = np.array(d["key_ctrl"], dtype=np.float64)
        self.initial_y = d["key_qpos"][1]; self.initial_z_flat = d["key_qpos"][2]
        self.n_sub = max(1, int(0.08 / d["timestep"]))                                           # :178, POLICY_DECISION_DT :106
        self.state_dim, self.action_dim = 12, 8
        self.steps = 0
        self.last_cmd = self.initial_ctrl.copy()
        self.prev_x = 0.0; self.cpos = 0.0; self.cneg = 0.0; self.prev_net = 0.0
        self.settled_z = self.initial_z_flat
        self.terrain = terrain
        self.hfield_data = np.zeros(TERRAIN_ROWS * TERRAIN_COLS, dtype=np.float32)

    def obs(self):                                                                               # :197-201
        s = self.sim
        yaw, pitch, roll = quat_to_ypr(s.qpos[3:7])
        jp = [s.qpos[q] - h for q, h in zip(self.qidx, self.home)]
        return np.concatenate([[yaw, pitch, roll], np.array(jp), [s.qvel[0]]]).astype(np.float32)

    def reset(self, seed=None):                                                                  # :321-346
        s = self.sim
        self.steps = 0
        if self.terrain:
            if seed is not None:
                _random.seed(seed)
            self.hfield_data = generate_terrain(_random, (s.desc["key_qpos"][0], s.desc["key_qpos"][1]))
        s.reset_keyframe()                       # mj_resetData in the generator, then qpos/qvel/ctrl = keyframe (:308-309)
        s.ctrl[:] = self.initial_ctrl
        s.forward()
        for _ in range(100):
            s.ctrl[:] = self.initial_ctrl
            s.step()
        s.forward()
        self.prev_x = float(s.qpos[0]); self.settled_z = float(s.qpos[2])
        self.last_cmd = s.ctrl.copy()
        self.cpos = self.cneg = self.prev_net = 0.0
        return self.obs()

    def step(self, a):                                                                           # :348-430
        s = self.sim
        self.steps += 1
        pol = np.asarray(a, dtype=np.float64) * AMP2
        cmd = np.zeros(8)
        for o in range(8):
            cmd[self.act_id[o]] = np.clip(self.home[o] + pol[o], self.ctrlrange[o][0], self.ctrlrange[o][1])
        s.ctrl[:] = cmd
        self.min_gap = np.full(3, 1e30)
        for _ in range(self.n_sub):
            s.step()
            self.min_gap = np.minimum(self.min_gap, s.decision_gaps)
        mj_err = not (np.isfinite(s.qpos).all() and np.isfinite(s.qvel).all())
        obs = self.obs()
        x = float(s.qpos[0]); dx = x - self.prev_x
        if dx > 0:
            self.cpos += dx
        elif dx < 0:
            self.cneg += abs(dx)
        self.prev_x = x
        vx = float(s.qvel[0])
        r_fwd = 450.0 * vx
        net = self.cpos - self.cneg; dnet = net - self.prev_net; self.prev_net = net
        r_prog = 20.0 * dnet if dnet > 0.0005 else 0.0
        p_back = -9.0 * abs(vx) if vx < -0.005 else 0.0
        r_disp = 0.0
        if dx > 0:
            r_disp = 70.0 * dx
        elif dx < 0.0005:
            r_disp = -1.0
        p_side = -0.3 * abs(s.qvel[1]); p_ypos = -0.15 * abs(s.qpos[1] - self.initial_y); p_yvel = -0.5 * abs(s.qvel[1])
        zs = s.qpos[2] - self.settled_z; zi = s.qpos[2] - self.initial_z_flat
        p_z = 0.0
        if zs < -0.03:
            p_z -= (0.25 * 0.5) * (abs(zs) - 0.03) ** 2
        if abs(zi) > 0.05:
            p_z -= (0.25 * 0.25) * (abs(zi) - 0.05) ** 2
        yaw, pitch, roll = quat_to_ypr(s.qpos[3:7])
        th, thy = math.radians(15.0), math.radians(35.0)
        p_or = 0.0
        if abs(roll) > th: p_or += -0.08 * (abs(roll) - th) ** 2
        if abs(pitch) > th: p_or += -0.08 * (abs(pitch) - th) ** 2
        if abs(yaw) > thy: p_or += -0.08 * (abs(yaw) - thy) ** 2
        ads = np.sum([(cmd[u] - self.last_cmd[u]) ** 2 for u in self.act_id])
        p_smooth = -0.005 * ads
        jvm = np.sum(np.abs(s.qvel[7:15]))              # (the reference's "joint velocities": qvel[7:15] = 7 of the 8 hinges)
        p_lowvel = -0.05 * np.exp(-jvm * 5.0)
        r = (r_fwd + r_prog + p_back + 0.005 + 0.01 + p_side + p_yvel + p_ypos + p_z + p_or + p_smooth + r_disp + p_lowvel)
        done, reason = False, "max_steps"
        lim = math.radians(35.0)
        if mj_err:
            r -= 50.0; done = True; reason = "mj_error"
        if not done and (abs(roll) > lim or abs(pitch) > lim or abs(yaw) > lim * 1.5):
            r -= 150; done = True; reason = "orientation_limit"
        if not done and self.cpos > 0.05 and self.cneg > 0.85 * self.cpos:
            r -= 50.0; done = True; reason = "too_much_backward"
        self.last_cmd = cmd.copy()
        return obs, float(r), done, {"sim_target_rad": cmd.copy(), "termination_reason": reason}
