#!/bin/bash
# A/B of library variants on the main bench path at 4096 and 65536 envs. usage: tools/gpu_quick5.sh variant...
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --mppi 0 --go1 0"
for v in default "$@"; do
  L=""; [ "$v" != default ] && L="ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so"
  for n in 4096 65536; do
    env $L $B --envs-per-gpu $n > gpurun_out/ab_${v}_$n.log 2>&1
    python - "gpurun_out/ab_${v}_$n.log" "$v" "$n" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('%-14s %6s envs: value %.3e  ms/step %.3f  e2e %.3e'%(sys.argv[2],sys.argv[3],d['value'],d['ms_per_step'],d['e2e']['value']))
except Exception as e: print('ERR',sys.argv[1],e, open(sys.argv[1]).read()[-300:])
PY
  done
done
