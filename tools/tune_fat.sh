#!/bin/bash
# Which instantiation / lockstep setting wins at the batch sizes between "one wave" and "many waves". Usage: tools/tune_fat.sh
for N in 8192 16384 32768; do for fat in 0 1; do for ls in 0 1; do
  python bench.py --steps 30 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N --cfg launch_fat=$fat --cfg launch_lockstep=$ls > gpurun_out/tune.log 2>&1
  python - $N $fat $ls <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tune.log").read().strip().splitlines()[-1]); print("envs %6s fat %s lockstep %s  value %.4e"%(*sys.argv[1:4], d["value"]), flush=True)
except Exception as e: print(sys.argv[1:4], "ERR", open("gpurun_out/tune.log").read()[-300:])
PY
done; done; done
