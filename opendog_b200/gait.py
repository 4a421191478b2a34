"""Gait files: the reference's real-robot interchange format and its sim<->real angle maps.

* JSON schema `[{"duration": seconds, "targets_deg": {actuator name: degrees}}]` — sim2real/walk.json,
  Code/examples/walks/*.json, written by `generate_walk_json` (sim2real/train.py:600-636) and consumed by the robot
  (`quadpilot/body.py` set_angles) and by the open-loop player sim2real/run.py:176-330.
* `export_walk_json`   = generate_walk_json: deterministic (mean-action) policy rollout -> real-robot degrees.
* `load_walk_json`     = run.py `load_and_process_sequence`: real degrees -> clamped sim ctrl targets.
* `play_walk`          = run.py `run_simulation` without the viewer: hold each target for its duration, one mj_step at a
                         time, for a whole batch of environments on the GPU.
"""
from __future__ import annotations

import json
import math

import numpy as np
import torch

# sim2real/train.py:95-103 (exporter) — scale factors all 1
TRAIN_REAL_HOME_DEG = {"FR_tigh_actuator": -45.0, "FR_knee_actuator": 45.0, "FL_tigh_actuator": 45.0, "FL_knee_actuator": 45.0,
                       "BR_tigh_actuator": 45.0, "BR_knee_actuator": -45.0, "BL_tigh_actuator": 45.0, "BL_knee_actuator": -45.0}
TRAIN_SCALE = {k: 1.0 for k in TRAIN_REAL_HOME_DEG}
# sim2real/run.py:25-37 (player)
RUN_REAL_HOME_DEG = dict(TRAIN_REAL_HOME_DEG)
RUN_SCALE = {"FR_tigh_actuator": 1.0, "FR_knee_actuator": 1.0, "BR_tigh_actuator": -1.0, "BR_knee_actuator": -1.0,
             "FL_tigh_actuator": -1.0, "FL_knee_actuator": 1.0, "BL_tigh_actuator": -1.0, "BL_knee_actuator": -1.0}
ACTUATOR_NAMES_ORDERED = ["FR_tigh_actuator", "FR_knee_actuator", "FL_tigh_actuator", "FL_knee_actuator",
                          "BR_tigh_actuator", "BR_knee_actuator", "BL_tigh_actuator", "BL_knee_actuator"]      # train.py:25-30


def sim_rad_to_real_deg(sim_rad, sim_home_rad, real_home_deg, scale=1.0):
    """convert_sim_rad_to_real_deg (sim2real/train.py:120-130)."""
    return real_home_deg + scale * math.degrees(sim_rad - sim_home_rad)


def real_deg_to_sim_rad(real_deg, sim_home_rad, real_home_deg, scale=1.0):
    """convert_real_deg_to_sim_rad (sim2real/run.py:60-79)."""
    if scale == 0:
        return None
    return sim_home_rad + math.radians(real_deg - real_home_deg) / scale


def load_walk_json(path_or_list, desc, real_home=RUN_REAL_HOME_DEG, scale=RUN_SCALE):
    """run.py `load_and_process_sequence`: -> (targets [S, nu] in ctrl order, NaN = actuator not mentioned in that
    step, durations [S]). Sim home = the keyframe's ctrl (run.py:126-152); targets are clamped to ctrlrange (:214-215)."""
    raw = path_or_list if isinstance(path_or_list, list) else json.load(open(path_or_list))
    names = list(desc["act_names"])
    home = {n: desc["key_ctrl"][i] for i, n in enumerate(names)}
    rng = {n: desc["act_ctrlrange"][i] for i, n in enumerate(names)}
    tg, du = [], []
    for step in raw:
        if not isinstance(step, dict) or "duration" not in step or "targets_deg" not in step:
            continue
        row = np.full(len(names), np.nan)
        any_ok = False
        for n, v in step["targets_deg"].items():
            if n not in home or not isinstance(v, (int, float)):
                continue
            r = real_deg_to_sim_rad(v, home[n], real_home[n], scale[n])
            if r is None:
                continue
            row[names.index(n)] = float(np.clip(r, rng[n][0], rng[n][1]))
            any_ok = True
        if not any_ok and len(step["targets_deg"]) > 0:
            continue
        tg.append(row); du.append(float(step["duration"]))
    return np.array(tg), np.array(du)


# sim2real/main.py:19-38 (the hand-coded trot's sim -> real map): the exporter's home table, scale factors all 1
MAIN_REAL_HOME_DEG = dict(TRAIN_REAL_HOME_DEG)
MAIN_SCALE = {k: 1.0 for k in MAIN_REAL_HOME_DEG}


def hand_coded_trot(desc, real_home=MAIN_REAL_HOME_DEG, scale=MAIN_SCALE, num_steps=12, phase_duration=0.40,
                    initial_hold=1.0, final_hold=1.0):
    """`create_control_sequence` (sim2real/main.py:63-151), the reference's open-loop trot: hold the keyframe's ctrl for
    1 s, then 12 phases of 0.4 s in which one diagonal pair (FR + BL on even phases, FL + BR on odd ones) swings — thigh
    +0.10 rad, knee flexed by -0.50 (front) / -0.35 (back) — while the other pair stands — thigh -0.10, knee extended by
    +0.15 (front) / +0.20 (back) —, every target clamped to the actuator's ctrlrange (:118-128), then 1 s at home: 6.8 s =
    3400 substeps of PD targets. Returns (targets_rad [14, nu] in ctrl order, targets_deg [14, nu] = the real-robot
    commands of main.py:130-136, durations [14])."""
    names = list(desc["act_names"])
    home = np.array(desc["key_ctrl"], dtype=np.float64)
    lo = np.array([r[0] for r in desc["act_ctrlrange"]]); hi = np.array([r[1] for r in desc["act_ctrlrange"]])
    idx = {n: names.index(n + "_actuator") for n in ("FR_tigh", "FR_knee", "BR_tigh", "BR_knee", "FL_tigh", "FL_knee", "BL_tigh", "BL_knee")}
    swing_thigh, stance_thigh = 0.10, -0.10
    swing_knee = {"F": -0.50, "B": -0.35}
    stance_knee = {"F": 0.15, "B": 0.2}
    rows, durs = [home.copy()], [initial_hold]
    for step in range(num_steps):
        swing = ("FR", "BL") if step % 2 == 0 else ("FL", "BR")
        row = home.copy()
        for leg in ("FR", "BR", "FL", "BL"):
            sw = leg in swing
            row[idx[leg + "_tigh"]] = home[idx[leg + "_tigh"]] + (swing_thigh if sw else stance_thigh)
            row[idx[leg + "_knee"]] = home[idx[leg + "_knee"]] + (swing_knee if sw else stance_knee)[leg[0]]
        rows.append(np.clip(row, lo, hi)); durs.append(phase_duration)
    rows.append(home.copy()); durs.append(final_hold)
    rad = np.array(rows)
    deg = np.array([[sim_rad_to_real_deg(float(r[u]), float(home[u]), real_home[n], scale[n]) for u, n in enumerate(names)] for r in rad])
    return rad, deg, np.array(durs)


def trot_json(desc, path=None, **kw):
    """The gait file main.py:236-252 writes for the robot: degrees rounded to 2 decimals, durations to 3."""
    rad, deg, durs = hand_coded_trot(desc, **kw)
    # (the reference's dicts are keyed in its ACTUATOR_NAMES order for the hold steps and in assignment order for the phases;
    #  consumers index by name)
    seq = [{"duration": round(float(d), 3), "targets_deg": {n: round(float(row[u]), 2) for u, n in enumerate(desc["act_names"])}}
           for row, d in zip(deg, durs)]
    if path:
        with open(path, "w") as f:
            json.dump(seq, f, indent=2)
    return seq


def segment_substeps(durations, timestep):
    """How many mj_step calls run.py spends on each sequence step: it advances when `data.time >= start + duration`,
    checked BEFORE each step with data.time accumulated in float64 (run.py:286-330)."""
    counts, t, start = [], 0.0, 0.0
    for d in durations:
        n = 0
        while not (t >= start + d):
            t += timestep
            n += 1
        counts.append(n)
        start = t
    return counts


def play_walk(env, targets, durations, initial_ctrl=None):
    """Open-loop playback on a BatchedWalkEnv created with scale_actions=0, auto_reset=0 (ctrl targets in rad).
    Every environment plays the same sequence from its current state. Returns qpos, qvel [N, .] at the end."""
    desc = env.desc
    cur = np.array(desc["key_ctrl"] if initial_ctrl is None else initial_ctrl, dtype=np.float64)
    counts = segment_substeps(durations, desc["timestep"])
    for row, n in zip(targets, counts):
        m = ~np.isnan(row)
        cur[m] = row[m]
        if n == 0:
            continue
        env.L.odg_set_frame_skip(env._h, int(n))
        a = torch.tensor(cur, dtype=torch.float32, device=env.device).reshape(1, -1).expand(env.num_envs, -1).contiguous()
        env.step_into(a, None, None, None, None)
    env.L.odg_set_frame_skip(env._h, int(env.cfg.frame_skip))
    return env.get_state()


def export_walk_json(agent, env, path, num_steps=50, real_home=TRAIN_REAL_HOME_DEG, scale=TRAIN_SCALE, env_index=0):
    """generate_walk_json (sim2real/train.py:600-636) on a BatchedQuadrupedEnv: mean-action rollout from reset, the
    commanded sim targets converted to real-robot degrees, rounded to 2 decimals, POLICY_DECISION_DT = 0.1 s per step."""
    desc = env.sim.desc
    names = list(desc["act_names"])
    # sim home of each actuator's JOINT = key_qpos (train.py:519-523)
    home = {n: desc["key_qpos"][7 + desc["act_leg"][i] * desc["njl"] + desc["act_joint"][i]] for i, n in enumerate(names)}
    seq = []
    obs = env.reset()
    for _ in range(num_steps):
        action, _, _, _ = agent.act(obs, sample=False)
        obs, _, done, info = env.step(action)
        cmd = info["sim_target_rad"][env_index].double().cpu().numpy()
        tdeg = {n: round(sim_rad_to_real_deg(float(cmd[names.index(n)]), home[n], real_home[n], scale[n]), 2)
                for n in ACTUATOR_NAMES_ORDERED}
        seq.append({"duration": round(0.10, 3), "targets_deg": tdeg})
        if bool(done[env_index]):
            break
    if path:
        with open(path, "w") as f:
            json.dump(seq, f, indent=2)
    return seq
