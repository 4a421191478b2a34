#!/bin/bash
# A/B of library variants (build/variants/libodgsim_<name>.so; "default" = the in-tree library) on the headline batch,
# twice each, plus Go1 and one MPPI plan. Usage: tools/ab_quick.sh name1 name2 ...
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = "default" ]; then unset ODG_LIB_PATH; else export ODG_LIB_PATH=$PWD/build/variants/libodgsim_$v.so; fi
    python bench.py --steps 40 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi $((rep == 2)) --go1 $((rep == 2)) > gpurun_out/ab.log 2>&1
    python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
    print("%-10s value %.4e e2e %.4e" % (sys.argv[1], d["value"], d["e2e"]["value"]), ("go1 %.4e mppi %.2f ms" % (d["step_go1"]["env_steps_per_s"], d["mppi"]["ms_per_plan"])) if "mppi" in d else "", flush=True)
except Exception as e: print(sys.argv[1], "ERR", open("gpurun_out/ab.log").read()[-300:])
PY
  done
done
