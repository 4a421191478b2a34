import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from opendog_b200.compat import BatchedQuadrupedEnv
from oracle.sim2real_oracle import QuadrupedEnvOracle
N,T=16,14
env=BatchedQuadrupedEnv(N,auto_reset=False); env.reset()
orcs=[QuadrupedEnvOracle() for _ in range(N)]; [o.reset() for o in orcs]
rng=np.random.default_rng(3); eo_all=[]; er_all=[]
for t in range(T):
    a=rng.uniform(-1,1,(N,4)).astype(np.float32)
    gq,gv=[x.cpu().numpy() for x in env.sim.get_state()]
    for i,o in enumerate(orcs): o.sim.qpos[:]=gq[i]; o.sim.qvel[:]=gv[i]
    obs,rew,done,info=env.step(torch.from_numpy(a)); obs=obs.cpu().numpy(); rew=rew.cpu().numpy()
    for i,o in enumerate(orcs):
        eo,er,ed,ei=o.step(a[i]); d=np.abs(obs[i]-eo); eo_all.append(d.max()); er_all.append(abs(rew[i]-er))
        if d.max()>2e-3: print(t,i,'obs err',d.max(),'idx',d.argmax(),'rew',rew[i],er)
eo_all=np.array(eo_all); er_all=np.array(er_all)
print('obs err pct 50/90/99/max',np.percentile(eo_all,[50,90,99,100])); print('rew err',np.percentile(er_all,[50,90,99,100]))
