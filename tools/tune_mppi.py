#!/usr/bin/env python
"""MPPI plan time (1024 samples x horizon 64, one fused rollout launch) against the launch shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.mppi import MPPI

def time_plan(**shape):
    m = MPPI(1024, 64, sigma=0.3, lam=1.0, seed=0, use_graph=True, **shape)
    q = torch.tensor(m.env.desc["key_qpos"], dtype=torch.float32); q[2] = 0.075
    m.set_start(q, torch.zeros(m.env.nv))
    for _ in range(3):
        m.plan()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        m.plan()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 5, float(m.stats[0])

for lanes, block in [(0, 0), (8, 64), (8, 128), (8, 32), (4, 32), (4, 64), (4, 128), (16, 64), (32, 64)]:
    ms, c = time_plan(launch_lanes=lanes, launch_block=block)
    print(f"lanes {lanes:2d} block {block:3d}: {ms:7.2f} ms per plan   (min cost {c:.3f})", flush=True)
