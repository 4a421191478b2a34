#!/usr/bin/env python
"""step_host with copies (0), with the actions read in place (1), with the results written in place too (2): same bits, time
per env-step. Usage: tools/ab_e2e.py [envs] [steps]"""
import os, sys, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendog_b200.env import BatchedWalkEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
with_info = len(sys.argv) > 3 and sys.argv[3] == "info"          # the SB3 adapter's call: all default info arrays too
g = torch.Generator().manual_seed(3)
acts = (torch.rand(steps + 22, n, 8, generator=g) * 2 - 1).pin_memory()
ref = None
for rep in range(2):
    for mode in (0, 1, 2):
        env = BatchedWalkEnv(n, seed=5, host_zero_copy=mode, **({} if with_info else dict(info_keys=None)))
        env.reset()
        for i in range(22):
            env.step_host(acts[i])
        gc.collect(); gc.disable()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(22, 22 + steps):
            o, r, te, tr = env.step_host(acts[i], with_info=with_info)[:4]
        e1.record(); torch.cuda.synchronize(); gc.enable()
        ms = e0.elapsed_time(e1) / steps
        out = (o.clone(), r.clone(), te.clone(), tr.clone())
        if ref is None: ref = out
        same = all(torch.equal(a, b) for a, b in zip(ref, out))
        print("envs %d mode %d: %.4f ms/step = %.4e env-steps/s  same bits as mode 0: %s" % (n, mode, ms, n / (ms * 1e-3), same), flush=True)
        env.close()
