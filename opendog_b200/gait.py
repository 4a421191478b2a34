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
