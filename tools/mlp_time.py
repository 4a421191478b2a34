#!/usr/bin/env python
"""Where one CTA of k_mlp spends its time: clock64 stamps of CTA 0's three roles (build with -DODG_MLP_TIMING).

    nvcc ... -DODG_MLP_TIMING -o build/variants/libodgsim_mlptime.so ;  python tools/mlp_time.py [rows]
"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['ODG_LIB_PATH'] = os.path.join(ROOT, 'build/variants/libodgsim_mlptime.so')
import torch
from opendog_b200.policy import ActorCriticB200
from opendog_b200 import lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = ActorCriticB200(33, 8, 0.4)
obs = torch.randn(N, 33, device='cuda')
for i in range(5): m.act(obs, step=i)
torch.cuda.synchronize()
L = lib.load(); buf = (C.c_longlong * 64)()
L.odg_mlp_timing(buf)
t = list(buf)
names = {0: 'epi: start'}
for net in range(2):
    b = net * 7
    names.update({1 + b: f'epi n{net} L1h0 acc ready', 2 + b: f'epi n{net} L1h0 drained', 3 + b: f'epi n{net} L1h1 acc ready', 4 + b: f'epi n{net} L1h1 drained',
                  5 + b: f'epi n{net} L2 acc ready', 6 + b: f'epi n{net} L2 drained', 7 + b: f'epi n{net} L3 acc ready'})
    b = 16 + net * 8
    names.update({b: f'mma n{net} L1h0 issued', b + 1: f'mma n{net} L1h1 issued', b + 2: f'mma n{net} L2 first half issued', b + 3: f'mma n{net} L2 issued', b + 4: f'mma n{net} L3 issued'})
names.update({15: 'epi: sampled', 32: 'tma: start', 33: 'tma n0 W1 sent', 34: 'tma n0 all sent', 35: 'tma n1 W1 sent', 36: 'tma: all sent'})
t0 = min(t[i] for i in names)
for i in sorted(names, key=lambda i: t[i]):
    print(f'{(t[i] - t0) / 1.965e3:8.2f} us  {names[i]}')
