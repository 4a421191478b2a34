#!/usr/bin/env python
"""Run the policy-forward kernel a few times on 16384 rows (profiling target: ncu -k regex:k_mlp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.policy import ActorCriticB200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = ActorCriticB200(33, 8, 0.4)
obs = torch.randn(N, 33, device="cuda")
for i in range(6):
    m.act(obs, step=i)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(20):
    m.act(obs, step=i)
e.record(); torch.cuda.synchronize()
print(f"k_mlp {N} rows: {s.elapsed_time(e) / 20 * 1e3:.1f} us per forward")
