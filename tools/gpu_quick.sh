#!/bin/bash
# quick perf visit: bench (default + variants) and one full ncu profile at 65536 envs
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
B="python bench.py --steps 10 --warmup 12 --no-cpu-baseline --large-batch 65536"
$B > gpurun_out/bench_default.log 2>&1
for v in "$@"; do
  ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so $B > gpurun_out/bench_$v.log 2>&1
done
Q="python bench.py --steps 3 --warmup 12 --no-cpu-baseline --large-batch 0 --envs-per-gpu 65536"
$Q > gpurun_out/plain_65536.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_65536 $Q > gpurun_out/ncu_f_65536.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.3f e2e %.3e large %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('large_batch',{}).get('env_steps_per_s')))
except Exception as e: print('  ERR',e, open(sys.argv[1]).read()[-400:])
PY
done
