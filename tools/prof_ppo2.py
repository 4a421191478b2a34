#!/usr/bin/env python
"""Eager vs CUDA-graph PPO epoch at the configs[3] size (65536 envs x 24 steps), per-call times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.policy import ActorCriticB200
from opendog_b200.train import ppo_update, GraphedPPOUpdate
B, S, A = 65536 * 24, 33, 8
dev = torch.device("cuda")
pol = ActorCriticB200(S, A, 0.4, device=dev, seed=0)
opt = torch.optim.Adam(pol.parameters(), lr=1e-4, fused=True, capturable=True)
obs = torch.randn(B, S, device=dev); act = torch.randn(B, A, device=dev).clamp(-1, 1)
logp = torch.randn(B, device=dev) * 0.1 - 8; adv = torch.randn(B, device=dev); ret = torch.randn(B, device=dev)
def timed(f, n=6):
    out = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); out.append(round(s.elapsed_time(e), 2))
    return out
print("eager", timed(lambda: ppo_update(pol, opt, obs, act, logp, adv, ret)), "MB", torch.cuda.max_memory_allocated() >> 20)
g = GraphedPPOUpdate(pol, opt, B, S, A)
print("graph", timed(lambda: g(obs, act, logp, adv, ret)), "MB", torch.cuda.max_memory_allocated() >> 20)
print("graph replay only", timed(lambda: [gr.replay() for gr in g.graphs]))
print({k: float(v) for k, v in g.out.items()})
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    [gr.replay() for gr in g.graphs]; torch.cuda.synchronize()
rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), reverse=True)
tot = sum(r[0] for r in rows)
print("kernel time of one epoch: %.3f ms in %d launches" % (tot / 1e3, sum(r[1] for r in rows)))
for t, n, k in rows[:70]:
    print("%9.1f us %5.1f %% %4d x  %s" % (t, 100 * t / tot, n, k[:120]))
