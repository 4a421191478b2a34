#!/usr/bin/env python
"""GPU diagnostic: find env-steps where the CUDA path and the oracle disagree beyond tolerance and replay
them substep by substep (state re-synchronised before every substep) to see which substep / contact set differs."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opendog_b200.env import BatchedWalkEnv
from oracle.oracle import WalkEnv, Sim, scale_action

N, T = 256, 40
env = BatchedWalkEnv(N, seed=11, max_episode_steps=25, info_keys=("ncon", "solver_iters"))
ws = [WalkEnv(seed=11, env_id=i) for i in range(N)]
for w in ws:
    w.e.max_steps = 25
env.reset(); [w.reset() for w in ws]
rng = np.random.default_rng(5)
sub = BatchedWalkEnv(1, frame_skip=1, scale_actions=0, auto_reset=0, info_keys=("ncon", "solver_iters", "contact_normal_force"))
sim = Sim()
nbad = 0
for t in range(T):
    a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
    pre = [(w.qpos.copy(), w.qvel.copy(), w.qacc_warmstart.copy()) for w in ws]
    obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
    obs = obs.cpu().numpy()
    res = [w.step_autoreset(a[i]) for i, w in enumerate(ws)]
    oobs = np.stack([r[0] for r in res])
    d = np.abs(obs - oobs).max(1)
    for i in np.nonzero(d > 2e-4)[0]:
        nbad += 1
        print(f"t={t} env={i} obs diff {d[i]:.3e} (idx {np.abs(obs[i]-oobs[i]).argmax()}) done={bool(done[i])}")
        q, v, wst = pre[i]
        ctrl = scale_action(a[i])
        sim.reset_keyframe(); sim.qpos[:] = q; sim.qvel[:] = v; sim.qacc_warmstart[:] = wst; sim.ctrl[:] = ctrl
        for s in range(10):
            sub.set_state(sim.qpos[None].astype(np.float32), sim.qvel[None].astype(np.float32), sim.qacc_warmstart[None].astype(np.float32))
            _, _, _, inf = sub.step(torch.from_numpy(ctrl[None]).cuda())
            sim.step()
            gq, gv = [x.cpu().numpy()[0] for x in sub.get_state()]
            fn = sum(c["force"][0] for c in sim.contacts())
            print(f"   sub {s}: dq {np.abs(gq-sim.qpos).max():.2e} dv {np.abs(gv-sim.qvel).max():.2e} ncon {int(inf['ncon'][0])}/{sim.ncon} "
                  f"iters {int(inf['solver_iters'][0])}/{sim.d.solver_iter} fn {float(inf['contact_normal_force'][0]):.4f}/{fn:.4f} "
                  f"min|dist-margin| {min([abs(c['dist']-0.001) for c in sim.contacts()] or [9]):.2e}")
    gq, gv = [x.cpu().numpy() for x in env.get_state()]
    for i, w in enumerate(ws):
        w.qpos[:] = gq[i]; w.qvel[:] = gv[i]
print("outliers", nbad, "of", N * T)
