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
