#!/usr/bin/env python
"""Tiny end-to-end exercise of every kernel (compute-sanitizer target): walk env, Go1, sim2real surface, policy, GAE, MPPI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.env import BatchedWalkEnv
from opendog_b200.compat import BatchedQuadrupedEnv
from opendog_b200.policy import ActorCriticB200
from opendog_b200.rollout import Rollout
from opendog_b200.mppi import MPPI
e = BatchedWalkEnv(70, seed=1, max_episode_steps=3)          # 70: exercises the padding environments
e.reset()
for t in range(5):
    e.step(torch.rand(70, 8, device="cuda") * 2 - 1)
g = BatchedWalkEnv(40, model="go1", seed=1, info_keys=None)
g.reset()
for t in range(2):
    g.step((torch.rand(40, 12, device="cuda") * 2 - 1) * 0.3)
q = BatchedQuadrupedEnv(20)
q.reset(); q.step(torch.zeros(20, 4))
pol = ActorCriticB200(e.obs_dim, e.act_dim, 0.4)
ro = Rollout(e, pol, horizon=3, use_graph=False)
ro.collect(); ro.advantages(group=False)
m = MPPI(64, 3, use_graph=False)
m.set_start(torch.tensor(m.env.desc["key_qpos"]), torch.zeros(m.env.nv)); m.plan()
torch.cuda.synchronize()
print("run_small ok")
