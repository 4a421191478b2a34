#!/bin/bash
# variant comparison at 65536 envs: usage tools/gpu_quick3.sh "variant[:ENV=VAL]"...
set -u
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 12 --no-cpu-baseline --large-batch 65536 --rollout-envs 0 --mppi 0 --go1 0"
$B > gpurun_out/bench_default.log 2>&1
for spec in "$@"; do
  v=${spec%%:*}; e=""; [ "$spec" != "$v" ] && e=${spec#*:}
  env $e ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so $B > gpurun_out/bench_${v}_${e}.log 2>&1
done
for f in gpurun_out/bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.3f e2e %.3e large %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('large_batch',{}).get('env_steps_per_s')))
except Exception as e: print('  ERR',e, open(sys.argv[1]).read()[-400:])
PY
done
