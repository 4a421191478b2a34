#!/usr/bin/env python
"""Distribution of Newton iterations / line-search passes of the last substep over a long random-action run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.env import BatchedWalkEnv
N = 4096
for model in ("our_robot", "go1"):
    env = BatchedWalkEnv(N, model=model, seed=0, info_keys=("solver_iters", "ls_evals", "ncon"))
    env.reset()
    it = []; ls = []; nc = []; nd = 0
    sc = 1.0
    for t in range(120):
        obs, r, d, info = env.step((torch.rand(N, env.act_dim, device="cuda") * 2 - 1) * sc)
        nd += int(d.sum())
        if t >= 12:
            it.append(info["solver_iters"].clone()); ls.append(info["ls_evals"].clone()); nc.append(info["ncon"].clone())
    it = torch.cat(it).float(); ls = torch.cat(ls).float(); nc = torch.cat(nc).float()
    q = lambda x, p: float(torch.quantile(x[:200000], p))
    print(f"{model}: iters mean {it.mean():.2f} p50 {q(it,.5):.0f} p90 {q(it,.9):.0f} p99 {q(it,.99):.0f} max {it.max():.0f} "
          f"frac>=30 {float((it>=30).float().mean()):.2e} | ls passes/iter {float(ls.sum()/it.sum()):.2f} | ncon mean {nc.mean():.1f} max {nc.max():.0f} "
          f"| episodes ended {nd} | obs finite {bool(torch.isfinite(obs).all())}")
