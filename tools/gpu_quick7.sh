#!/bin/bash
set -u
B="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --mppi 0 --go1 0"
for cfg in "" "--cfg ls_iterations=3" "--cfg ls_iterations=5" "--cfg ls_tolerance=0.3" "--cfg ls_tolerance=0.03" "--cfg solver_tolerance=1e-4"; do
  for n in 4096 65536; do
    $B --envs-per-gpu $n $cfg | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs $n [$cfg]: %.3e'%d['value'])"
  done
done
