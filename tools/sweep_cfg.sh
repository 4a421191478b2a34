#!/bin/bash
# Throughput of the step against solver settings (OdgEnvConfig overrides). Usage: tools/sweep_cfg.sh "k=v k=v" "k=v" ...
for N in 4096 65536; do
  for cfg in "$@"; do
    args=""; for kv in $cfg; do [ "$kv" = "-" ] || args="$args --cfg $kv"; done
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N $args > gpurun_out/sweep.log 2>&1
    python - "$N" "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/sweep.log").read().strip().splitlines()[-1]); print("envs %6s  %-50s value %.4e"%(sys.argv[1], sys.argv[2], d["value"]), flush=True)
except Exception as e: print(sys.argv[1:3], "ERR", open("gpurun_out/sweep.log").read()[-300:])
PY
  done
done
