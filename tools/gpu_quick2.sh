#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
B="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 0 --regroup 0"
for l in 32 16 8; do
  ODG_STEP_LANES=$l $B > gpurun_out/bench_lanes$l.log 2>&1
  ODG_STEP_LANES=$l $B --cfg ls_tolerance=0.3 > gpurun_out/bench_lanes${l}_lstol03.log 2>&1
done
ODG_STEP_LANES=16 $B --envs-per-gpu 16384 > gpurun_out/bench_16k_lanes16.log 2>&1
ODG_STEP_LANES=32 $B --envs-per-gpu 16384 > gpurun_out/bench_16k_lanes32.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.3f e2e %.3e large %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('large_batch',{}).get('env_steps_per_s')))
except Exception as e: print('  ERR',e, open(sys.argv[1]).read()[-400:])
PY
done
