#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full profiles (each ncu pass only after
# the same command exited 0 without ncu). Usage: tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_default.log 2>&1
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_reference.log 2>&1
P="python bench.py --steps 3 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --mppi 0 --go1 0"
$P > gpurun_out/plain_4096.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_4096_$TAG.csv $P > gpurun_out/ncu_l_4096.log 2>&1
$P > gpurun_out/plain_4096b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_4096_$TAG $P > gpurun_out/ncu_f_4096.log 2>&1
Q="$P --envs-per-gpu 65536"
$Q > gpurun_out/plain_65536.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_65536_$TAG $Q > gpurun_out/ncu_f_65536.log 2>&1
M="python tools/run_mlp.py 16384"
$M > gpurun_out/plain_mlp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mlp -s 6 -c 1 -o gpurun_out/prof_mlp_$TAG $M > gpurun_out/ncu_f_mlp.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/plain_mlp.log
for f in gpurun_out/bench_*.log gpurun_out/plain_4*.log gpurun_out/plain_6*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.3f e2e %.3e large %s cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('large_batch',{}).get('env_steps_per_s'),d.get('cpu_baseline',{}).get('value')))
    for k in ('rollout','mppi'):
        if k in d: print('  ',k,{a:b for a,b in d[k].items() if not isinstance(b,(dict,str))})
except Exception as e: print('  ERR',e, open(sys.argv[1]).read()[-600:])
PY
done
