#!/bin/bash
# round-2 first GPU pass: the whole GPU test suite, the default bench line, launch-shape matrices
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r02a_pytest.log
python bench.py --steps 30 --warmup 12 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02a_bench.json").read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e large %.4e mppi %.2f ms rollout %.4e go1 %.4e cpu %.4e" % (d["value"], d["e2e"]["value"], d["large_batch"]["env_steps_per_s"],
          d["mppi"]["ms_per_plan"], d["rollout"]["env_steps_per_s_rollout"], d["step_go1"]["env_steps_per_s"], d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench parse error", e); print(open("gpurun_out/r02a_bench.err").read()[-2000:])
PY
python tools/tune_mppi.py 2>&1 | tee gpurun_out/r02a_mppi.txt
bash tools/tune_launch_shape.sh 4096 2>&1 | tee gpurun_out/r02a_shape_4096.txt
