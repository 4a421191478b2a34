import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
os.environ['ODG_LIB_PATH'] = '/root/repo/build/variants/libodgsim_mlptime.so'
import torch
from opendog_b200.policy import ActorCriticB200
from opendog_b200 import lib
m = ActorCriticB200(33, 8, 0.4)
obs = torch.randn(16384, 33, device='cuda')
for i in range(5): m.act(obs, step=i)
torch.cuda.synchronize()
L = lib.load(); buf = (C.c_longlong * 64)()
L.odg_mlp_timing(buf)
t = list(buf); t0 = t[0]
names = {0:'setup done'}
for net in range(2):
    b = net*12
    names.update({1+b:f'n{net} L1h0 acc ready', 2+b:f'n{net} L1h0 epilogue done', 3+b:f'n{net} L1h1 acc ready', 4+b:f'n{net} L1h1 epilogue done', 5+b:f'n{net} L2 acc ready', 6+b:f'n{net} L2 epilogue done', 7+b:f'n{net} L3 acc ready'})
names[30]='before sampling'
prev=t0
for i in sorted(names):
    print(f'{names[i]:28s} +{(t[i]-prev)/1.965e3:7.2f} us   (t={(t[i]-t0)/1.965e3:7.2f})'); prev=t[i]
