"""Test infrastructure (oracle side): probes for the MuJoCo pin built from the reference's shipped (checkpoint, walk file)
pairs — see tools/pin_walk_json.py for what the pin is and tests/test_mujoco_pin_walk_json.py for the gate. Follows
/root/reference/Code/mujoco/sim2real/train.py: ActorCritic.actor (:132-149), the policy -> target mapping of
`_apply_actions_and_step` (:235-277) and `convert_sim_rad_to_real_deg` (:120-130), reset (:209-233), `generate_walk_json`
(:600-636). Only tests/ and tools/ import this module."""
import copy

import numpy as np

from opendog_b200.model.compile import load_compiled

from .oracle import Sim
from .sim2real_oracle import AMP, QuadrupedEnvOracle

MATCHING = (2600, 2700, 2800, 2900, 3000, 3100, 3200, 3300, 3500, 3700)    # step 0 within 0.5 degree (section 1 of the report)
WIDE = dict(thigh_hi=3.2, knee_lo=-2.4)                                    # any range the 40-degree amplitude cannot reach
REAL_HOME_DEG = np.array([-45.0, 45.0, 45.0, 45.0, 45.0, -45.0, 45.0, -45.0])   # train.py:95-101 in ORDERED order


def actor_forward(w, x):
    """The reference ActorCritic's actor (train.py:132-149: Linear-Tanh-Linear-Tanh-Linear-Tanh), float64 numpy."""
    h = np.tanh(w[0] @ x + w[1]); h = np.tanh(w[2] @ h + w[3])
    return np.tanh(w[4] @ h + w[5])


def actor_weights(state_dict):
    return [np.asarray(state_dict[k], dtype=np.float64) for k in
            ("actor.0.weight", "actor.0.bias", "actor.2.weight", "actor.2.bias", "actor.4.weight", "actor.4.bias")]


def model_desc(wide):
    d = copy.deepcopy(load_compiled("our_robot"))
    if wide:
        for leg in range(4):
            d["jnt_range"][leg][0] = [2.36, WIDE["thigh_hi"]]; d["jnt_range"][leg][1] = [WIDE["knee_lo"], -1.2]
        for u in range(8):
            d["act_ctrlrange"][u] = [2.36, WIDE["thigh_hi"]] if d["act_joint"][u] == 0 else [WIDE["knee_lo"], -1.2]
    return d


class Probe:
    """QuadrupedEnvOracle on a given model description, with the policy -> targets mapping of train.py:235-277, 120-130."""

    def __init__(self, desc):
        self.e = QuadrupedEnvOracle(); self.e.sim = Sim(desc)
        self.home = np.array(self.e.home)
        self.cr = np.array([desc["act_ctrlrange"][u] for u in self.e.act_id])

    def reset(self, settle=100):
        e = self.e; s = e.sim
        e.counter = 0; s.reset_keyframe(); s.ctrl[:] = e.initial_ctrl
        for _ in range(settle):
            s.ctrl[:] = e.initial_ctrl; s.step()
        return e.obs().astype(np.float64)

    def targets_deg(self, w, x):
        fr, k1, fl, k2 = actor_forward(w, x) * AMP
        d = np.array([fr, 0, fl, 0, fl, 0, fr, 0.0])
        if self.e.counter % 2 == 0:
            d[1], d[7] = k1, -k1
        else:
            d[3], d[5] = k2, -k2
        return REAL_HOME_DEG + np.degrees(np.clip(self.home + d, self.cr[:, 0], self.cr[:, 1]) - self.home)

    def apply_deg(self, deg):
        e = self.e; cmd = np.zeros(8)
        for o in range(8):
            cmd[e.act_id[o]] = self.home[o] + np.radians(deg[o] - REAL_HOME_DEG[o])
        e.sim.ctrl[:] = cmd
        for _ in range(e.n_sub):
            e.sim.step()
        e.counter += 1
        return e.obs().astype(np.float64)


def teacher_forced(desc, w, shipped, nsteps):
    p = Probe(desc); x = p.reset(); err = []
    for t in range(min(nsteps, len(shipped))):
        err.append(np.abs(np.round(p.targets_deg(w, x), 2) - shipped[t]).max())
        x = p.apply_deg(shipped[t])
    return np.array(err)


def noise_floor(desc, w, nsteps):
    p = Probe(desc); x = p.reset(); rec = []
    for _ in range(nsteps):
        m = p.targets_deg(w, x); rec.append(np.round(m, 2)); x = p.apply_deg(m)
    return teacher_forced(desc, w, np.array(rec), nsteps)


class KernelProbe:
    """The same probe on the PRODUCT's physics: an environment object with the walk-environment surface (`reset()`,
    `step(ctrl)`, `get_state()`) created with `scale_actions=0, frame_skip=50, reset_noise_scale=0, auto_reset=0` — the
    CUDA kernel through the C ABI (opendog_b200.env.BatchedWalkEnv) or its source run by the host lane emulator
    (tests/emu) — stepped with the shipped targets as raw controls; the QuadrupedEnv observation (train.py:184-207) is
    rebuilt from qpos / qvel. `to_np` converts what get_state returns to numpy, `ctrl_of` a numpy row to what step takes."""

    def __init__(self, env, desc, to_np=np.asarray, ctrl_of=lambda a: a):
        from .sim2real_oracle import quat_to_ypr
        self.env, self.to_np, self.ctrl_of, self.ypr = env, to_np, ctrl_of, quat_to_ypr
        ref = QuadrupedEnvOracle()                          # (index maps of the ordered actuators)
        self.act_id, self.qidx, self.vidx = ref.act_id, ref.qidx, ref.vidx
        self.home = np.array(ref.home)
        self.cr = np.array([desc["act_ctrlrange"][u] for u in ref.act_id])
        self.key_ctrl = np.array(desc["key_ctrl"], np.float32)
        self.counter = 0

    def obs(self):
        st = self.env.get_state()
        qpos, qvel = self.to_np(st[0])[0].astype(np.float64), self.to_np(st[1])[0].astype(np.float64)
        yaw, pitch, roll = self.ypr(qpos[3:7])
        jp = [qpos[q] - h for q, h in zip(self.qidx, self.home)]
        jv = [qvel[v] for v in self.vidx]
        ph = (self.counter % 2) / 1.0
        return np.concatenate([[yaw, pitch, roll], jp, jv, [qvel[0]],
                               [np.sin(ph * np.pi), np.cos(ph * np.pi)]]).astype(np.float32).astype(np.float64)

    def reset(self):
        self.env.reset(); self.counter = 0
        for _ in range(2):                                  # 100 settling substeps with the keyframe's controls (train.py:218-221)
            self.env.step(self.ctrl_of(self.key_ctrl[None, :]))
        return self.obs()

    targets_deg = Probe.targets_deg

    @property
    def e(self):                                            # (Probe.targets_deg reads self.e.counter)
        return self

    def apply_deg(self, deg):
        cmd = np.zeros(8, np.float32)
        for o in range(8):
            cmd[self.act_id[o]] = self.home[o] + np.radians(deg[o] - REAL_HOME_DEG[o])
        self.env.step(self.ctrl_of(cmd[None, :])); self.counter += 1
        return self.obs()


def teacher_forced_kernel(probe, w, shipped, nsteps):
    x = probe.reset(); err = []
    for t in range(min(nsteps, len(shipped))):
        err.append(np.abs(np.round(probe.targets_deg(w, x), 2) - shipped[t]).max())
        x = probe.apply_deg(shipped[t])
    return np.array(err)
