#!/usr/bin/env python
"""Step the 12-actuator Go1 model a few times (profiling / tuning target). Usage: tools/run_go1.py [envs] [steps] [launch_x=v ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.env import BatchedWalkEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cfg = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:])}          # launch_* overrides: key=value ...
env = BatchedWalkEnv(n, model="go1", seed=0, info_keys=None, **cfg)
env.reset()
acts = torch.rand(steps, n, env.act_dim, device="cuda") * 2 - 1
for i in range(steps - 4):
    env.step(acts[i])
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(steps - 4, steps):
    env.step(acts[i])
e.record(); torch.cuda.synchronize()
print("go1 %d envs %s: %.4f ms/step = %.4e env-steps/s" % (n, cfg, s.elapsed_time(e) / 4, n / (s.elapsed_time(e) / 4 * 1e-3)))
