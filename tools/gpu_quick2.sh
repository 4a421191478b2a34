#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python tools/diag_outliers.py > gpurun_out/diag.log 2>&1
B="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 65536"
$B --regroup 0 > gpurun_out/bench_rg0.log 2>&1
$B --regroup 1 > gpurun_out/bench_rg1.log 2>&1
python bench.py --steps 20 --warmup 40 --no-cpu-baseline --large-batch 131072 --regroup 1 > gpurun_out/bench_rg1_w40.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -25 gpurun_out/diag.log
for f in gpurun_out/bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.3f e2e %.3e large %s launches %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('large_batch',{}),d['gpu_launches']))
except Exception as e: print('  ERR',e, open(sys.argv[1]).read()[-400:])
PY
done
