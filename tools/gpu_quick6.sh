#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --mppi 0 --go1 0"
for n in 4096 8192 16384 32768 65536; do
  for ls in 0 1; do
    ODG_LOCKSTEP=$ls $B --envs-per-gpu $n | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs $n lockstep $ls: %.3e'%d['value'])"
  done
done
