#!/bin/bash
# Launch-shape matrix at a given batch size: environments per warp x block size x lockstep.
# Usage: tools/tune_launch_shape.sh [envs]
N=${1:-4096}
P="python bench.py --steps 40 --warmup 10 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N"
for lanes in 32 16; do for blk in 32 64 128; do for ls in 0 1; do
  $P --cfg launch_lanes=$lanes --cfg launch_block=$blk --cfg launch_lockstep=$ls > gpurun_out/tune.log 2>&1
  python - $lanes $blk $ls <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tune.log").read().strip().splitlines()[-1]); print("lanes %s block %s lockstep %s  value %.4e"%(*sys.argv[1:4], d["value"]))
except Exception as e: print(sys.argv[1:4], "ERR", open("gpurun_out/tune.log").read()[-300:])
PY
done; done; done
