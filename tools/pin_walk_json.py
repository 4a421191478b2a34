#!/usr/bin/env python
"""A MuJoCo pin for the physics oracle, from artefacts the reference itself ships.

For every 100th training episode N the reference's trainer saved the policy it had
(`sim2real/output/pth/quadruped_ac_sym_epN.pth`, sim2real/train.py:587-588) and, in the same breath, the 50-step
deterministic closed-loop walk that policy produced in REAL MuJoCo 3.2.3 (`sim2real/output/json/walk_rl_sym_epN.json`,
written by `generate_walk_json`, train.py:589, 600-636: reset = home keyframe + 100 settling `mj_step`s, action = the
policy's mean on the observation of `QuadrupedEnv`, 50 `mj_step`s per action, commanded targets rounded to 0.01 degree).
Every recorded target is the policy's reading of a state that real MuJoCo computed: target t depends on 100 + 50 t
engine steps. The policies saturate easily (high gain), so they are sensitive probes of that state.

What this script does with them (all on the CPU, oracle physics = oracle/odg_oracle.c):

  1. closed loop: the reference's `generate_walk_json` UNMODIFIED, imported from /root/reference, `mujoco` stubbed onto
     the oracle, with every shipped checkpoint -> compared with the shipped file (which episodes were produced by the
     revision of the environment code that is in the tree: step 0 depends on the reset state and the code only);
  2. the settling transient: the step-0 targets as a function of the number of settling steps (the reference: 100);
  3. teacher forcing: the SHIPPED targets are applied as controls (so that both engines see the same control sequence and
     the policy's gain does not amplify differences), and at every step the policy's output on the oracle's observation is
     compared with the shipped target — with the joint / control ranges of the tree's `our_robot.xml` and with wider ones
     (the shipped targets exceed the tree's knee and thigh ranges: the XML was revised after the files were written);
  4. the noise floor of (3): the oracle against its OWN rounded closed-loop record (0.01-degree rounding alone).

    python tools/pin_walk_json.py [--write]   # --write: tests/golden/mujoco_pin_walk_json.npz + profiles/r02g_mujoco_pin.txt
"""
import builtins
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/mujoco"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle.sim2real_oracle import ORDERED  # noqa: E402

from oracle.mujoco_pin import (MATCHING, REAL_HOME_DEG, WIDE, Probe, actor_weights, model_desc, noise_floor,  # noqa: E402
                               teacher_forced)


def closed_loop_with_reference_code(eps):
    """Section 1: the reference's own generate_walk_json on the stubbed engine."""
    import torch
    from make_golden_sim2real import install_mujoco_stub
    install_mujoco_stub()
    sys.path.insert(0, os.path.join(REF, "sim2real"))
    real_print = builtins.print
    builtins.print = lambda *a, **k: None
    try:
        import train as ref
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
        assert list(ref.ACTUATOR_NAMES_ORDERED) == list(ORDERED)
        assert np.allclose([ref.real_robot_home_deg_map[n] for n in ORDERED], REAL_HOME_DEG)
        env = ref.QuadrupedEnv("our_robot/walking_scene.xml")
        rows = []
        for ep in eps:
            agent = ref.ActorCritic(env.state_dim, env.action_dim, ref.INITIAL_ACTION_STD_INIT)
            agent.load_state_dict(torch.load(os.path.join(REF, f"sim2real/output/pth/quadruped_ac_sym_ep{ep}.pth"),
                                             map_location="cpu", weights_only=True))
            with tempfile.TemporaryDirectory() as td:
                p = os.path.join(td, "walk.json")
                ref.generate_walk_json(agent, env, torch.device("cpu"), p, ref.JSON_MAX_STEPS_EPISODIC)
                mine = json.load(open(p)) if os.path.exists(p) else []
            a = shipped_targets(ep)
            b = np.array([[s["targets_deg"][n] for n in ORDERED] for s in mine])
            n = min(len(a), len(b))
            rows.append((ep, len(a), len(b), np.abs(a[:n] - b[:n]).max(axis=1)))
    finally:
        builtins.print = real_print
    return rows


def shipped_targets(ep):
    s = json.load(open(os.path.join(REF, f"sim2real/output/json/walk_rl_sym_ep{ep}.json")))
    return np.array([[st["targets_deg"][n] for n in ORDERED] for st in s])


def main():
    import torch
    out = io.StringIO()

    def say(*a):
        print(*a); print(*a, file=out)
    np.set_printoptions(precision=2, suppress=True, linewidth=220)
    all_eps = sorted(int(f[len("walk_rl_sym_ep"):-5]) for f in os.listdir(os.path.join(REF, "sim2real/output/json"))
                     if f[len("walk_rl_sym_ep"):-5].isdigit()
                     and os.path.exists(os.path.join(REF, f"sim2real/output/pth/quadruped_ac_sym_ep{f[len('walk_rl_sym_ep'):-5]}.pth")))
    say("# MuJoCo pin from the reference's shipped (checkpoint, walk file) pairs — tools/pin_walk_json.py")
    say("\n## 1. closed loop, the reference's generate_walk_json on the oracle's physics vs the shipped file (max |d| over the 8 targets, degrees)")
    say("episode  steps shipped/oracle   step 0      1      2      5     10")
    for ep, la, lb, d in closed_loop_with_reference_code(all_eps):
        pick = [d[i] if i < len(d) else float("nan") for i in (0, 1, 2, 5, 10)]
        say(f"{ep:6d}      {la:3d}/{lb:3d}         " + " ".join(f"{x:6.2f}" for x in pick) + ("   <- same code revision" if ep in MATCHING else ""))
    say("(step 0 depends on the code and the settled reset state only: episodes <= 2500 were written by an older revision of the"
        " action mapping; the closed loop decorrelates after one step everywhere — see 3 for why)")
    W = {ep: actor_weights(torch.load(os.path.join(REF, f"sim2real/output/pth/quadruped_ac_sym_ep{ep}.pth"), map_location="cpu",
                                      weights_only=True)) for ep in MATCHING}
    S = {ep: shipped_targets(ep) for ep in MATCHING}
    tree, wide = model_desc(False), model_desc(True)
    say("\n## 2. step-0 targets vs the number of settling mj_steps before the first observation (reference: 100); rms / max over the"
        " unclipped targets of the 10 matching episodes, degrees")
    for n in (95, 98, 99, 100, 101, 102, 105):
        r = []
        for ep in MATCHING:
            p = Probe(tree); x = p.reset(n); t = p.targets_deg(W[ep], x)
            r += [t[i] - S[ep][0][i] for i in (1, 2) if p.cr[i, 0] < p.home[i] + np.radians(t[i] - REAL_HOME_DEG[i]) < p.cr[i, 1]]
        r = np.array(r)
        say(f"  settle {n:3d}: rms {np.sqrt((r ** 2).mean()):6.3f}  max {np.abs(r).max():6.3f}   ({len(r)} targets)")
    N = 12
    say(f"\n## 3. teacher forcing with the shipped targets, max |d| per step (first {N} steps), degrees")
    say("with the tree's ranges (thigh [2.36, 2.8], knee [-1.8, -1.2]):")
    Et = np.array([np.pad(teacher_forced(tree, W[ep], S[ep], N), (0, max(0, N - len(S[ep]))), constant_values=np.nan) for ep in MATCHING])
    for ep, e in zip(MATCHING, Et):
        say(f"  ep {ep}: {e}")
    say(f"with wide ranges (thigh [2.36, {WIDE['thigh_hi']}], knee [{WIDE['knee_lo']}, -1.2]; shipped knee targets reach -85 deg = home - 40,"
        " shipped thigh targets 84.7 = home + 39.5: beyond the tree's ranges):")
    Ew = np.array([np.pad(teacher_forced(wide, W[ep], S[ep], N), (0, max(0, N - len(S[ep]))), constant_values=np.nan) for ep in MATCHING])
    for ep, e in zip(MATCHING, Ew):
        say(f"  ep {ep}: {e}")
    say(f"  median over episodes, tree ranges: {np.nanmedian(Et, axis=0)}")
    say(f"  median over episodes, wide ranges: {np.nanmedian(Ew, axis=0)}")
    say("\n## 4. noise floor: the oracle against its own closed-loop record rounded to 0.01 degree (wide ranges)")
    F = np.array([noise_floor(wide, W[ep], N) for ep in MATCHING])
    say(f"  median over episodes: {np.median(F, axis=0)}")
    say(f"  max over episodes:    {F.max(axis=0)}")
    if "--write" in sys.argv:
        keep = {f"shipped_{ep}": S[ep] for ep in MATCHING}
        for ep in (2700, 3200):                                   # (ep 3700's checkpoint is already a fixture)
            for i, a in enumerate(W[ep]):
                keep[f"actor_{ep}_{i}"] = a.astype(np.float32)
        keep["teacher_forced_wide"] = Ew; keep["teacher_forced_tree"] = Et; keep["episodes"] = np.array(MATCHING)
        path = os.path.join(ROOT, "tests", "golden", "mujoco_pin_walk_json.npz")
        np.savez_compressed(path, **keep)
        with open(os.path.join(ROOT, "profiles", "r02g_mujoco_pin.txt"), "w") as f:
            f.write(out.getvalue())
        print("wrote", path, "and profiles/r02g_mujoco_pin.txt")


if __name__ == "__main__":
    main()
