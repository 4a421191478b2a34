#!/bin/bash
# quick perf visit: parity tests + bench at 4096 / 65536 envs for the default build, config overrides and variants
# usage: tools/gpu_quick.sh [variant ...]   (variants: build/variants/libodgsim_<v>.so)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
B="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 65536"
$B --regroup 0 > gpurun_out/bench_default_rg0.log 2>&1
$B --regroup 1 > gpurun_out/bench_default_rg1.log 2>&1
$B --regroup 0 --cfg ls_tolerance=0.1 > gpurun_out/bench_lstol01.log 2>&1
$B --regroup 0 --cfg ls_tolerance=0.3 > gpurun_out/bench_lstol03.log 2>&1
for v in "$@"; do
  ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so $B --regroup 0 > gpurun_out/bench_$v.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.3f e2e %.3e large %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('large_batch',{}).get('env_steps_per_s')))
except Exception as e: print('  ERR',e, open(sys.argv[1]).read()[-400:])
PY
done
