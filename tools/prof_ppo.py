#!/usr/bin/env python
"""Where the PPO epoch's time goes (torch profiler, synthetic rollout of 65536 envs x 24 steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.policy import ActorCriticB200
from opendog_b200.train import ppo_update
B, S, A = 65536 * 24, 33, 8
dev = torch.device("cuda")
pol = ActorCriticB200(S, A, 0.4, device=dev, seed=0)
opt = torch.optim.Adam(pol.parameters(), lr=1e-4, fused=True)
obs = torch.randn(B, S, device=dev); act = torch.randn(B, A, device=dev).clamp(-1, 1)
logp = torch.randn(B, device=dev) - 8; adv = torch.randn(B, device=dev); ret = torch.randn(B, device=dev)
for ac in (False, True):
    for _ in range(2):
        ppo_update(pol, opt, obs, act, logp, adv, ret, autocast=ac)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ppo_update(pol, opt, obs, act, logp, adv, ret, autocast=ac); e.record(); torch.cuda.synchronize()
    print(f"autocast={ac}: {s.elapsed_time(e):.2f} ms per epoch")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        ppo_update(pol, opt, obs, act, logp, adv, ret, autocast=ac)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
# the padded bf16 forward must be the reference module's forward (fp32) up to bf16 rounding
with torch.no_grad():
    x = obs[:4096]
    d0, v0 = pol(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        d1, v1 = pol.forward_padded(pol.pad_obs(x.to(torch.bfloat16)))
    print("padded bf16 vs fp32 forward: mean", (d0.mean - d1.mean.float()).abs().max().item(), "value", (v0 - v1.float()).abs().max().item())
