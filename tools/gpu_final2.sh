#!/bin/bash
# Round-end confirmation of the final launch defaults: GPU suite, smoke, default bench, reference arm, Go1 / FAT spot checks.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02g_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02g_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02g_smoke.log
python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02g_reference.json 2> gpurun_out/r02g_reference.err; echo "reference exit $?"
for N in 4096 16384 65536; do python tools/run_go1.py $N 28; done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d["roofline"])
for k in ("large_batch","rollout","configs3","go1","mppi","cpu_baseline","clocks"):
    if k in d: print(k, json.dumps(d[k])[:700])
r=json.loads(open("gpurun_out/r02g_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r.get("cpu_baseline"))
PY
