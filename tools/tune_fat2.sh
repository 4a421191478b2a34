#!/bin/bash
# FAT vs lean instantiation (x lockstep mode) at the batch sizes between one wave and many. Usage: tools/tune_fat2.sh
mkdir -p gpurun_out
for N in 6144 8192 12288 16384 32768 65536; do for cfg in "0 2" "0 1" "1 2" "1 0" "1 1"; do
  set -- $cfg
  python bench.py --steps 30 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N --cfg launch_fat=$1 --cfg launch_lockstep=$2 > gpurun_out/tune.log 2>&1
  python - $N $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tune.log").read().strip().splitlines()[-1]); print("envs %6s fat %s lockstep %s  value %.4e"%(*sys.argv[1:4], d["value"]), flush=True)
except Exception as e: print(sys.argv[1:4], "ERR", open("gpurun_out/tune.log").read()[-300:])
PY
done
python bench.py --steps 30 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N > gpurun_out/tune.log 2>&1
python - $N <<'PY'
import json,sys
d=json.loads(open("gpurun_out/tune.log").read().strip().splitlines()[-1]); print("envs %6s auto                value %.4e"%(sys.argv[1], d["value"]), flush=True)
PY
done
