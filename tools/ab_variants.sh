#!/bin/bash
# A/B of library variants (build/variants/*.so) at 4096 and 65536 envs. Usage: tools/ab_variants.sh name1 name2 ...
for v in "$@"; do
  if [ "$v" = default ]; then unset ODG_LIB_PATH; else export ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so; fi
  for cfg in "4096 32 0" "65536 32 1"; do
    set -- $cfg
    python bench.py --steps 30 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --mppi 0 --go1 0 --envs-per-gpu $1 --cfg launch_lanes=$2 --cfg launch_lockstep=$3 > gpurun_out/q10.log 2>&1
    python - "$v" $cfg <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/q10.log").read().strip().splitlines()[-1]); print("%-10s envs %s lanes %s lockstep %s  value %.4e"%(*sys.argv[1:5], d["value"]))
except Exception as e: print(sys.argv[1:5], "ERR", open("gpurun_out/q10.log").read()[-300:])
PY
  done
done
