#!/bin/bash
# A/B of library variants (build/variants/*.so, ODG_LIB_PATH) at 4096 and 65536 envs. Usage: tools/ab_libs.sh name1 name2 ...
for N in 4096 65536; do
  for v in "$@"; do
    if [ "$v" = "default" ]; then unset ODG_LIB_PATH; else export ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so; fi
    python bench.py --steps 40 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N > gpurun_out/ab.log 2>&1
    python - "$N" "$v" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1]); print("envs %6s  %-12s value %.4e"%(sys.argv[1], sys.argv[2], d["value"]), flush=True)
except Exception as e: print(sys.argv[1:3], "ERR", open("gpurun_out/ab.log").read()[-300:])
PY
  done
done
