#!/bin/bash
# MPPI (1024 samples x 64) against environments per warp / block size.
P="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --go1 0"
for lanes in 32 16 8; do for blk in 32 64; do
  ODG_STEP_LANES=$lanes ODG_STEP_BLOCK=$blk $P > gpurun_out/q9.log 2>&1
  python - $lanes $blk <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/q9.log").read().strip().splitlines()[-1]); print("lanes %s block %s  ms_per_plan %.2f"%(*sys.argv[1:3], d["mppi"]["ms_per_plan"]))
except Exception as e: print(sys.argv[1:3], "ERR", open("gpurun_out/q9.log").read()[-300:])
PY
done; done
