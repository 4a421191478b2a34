#!/bin/bash
# round-2 GPU pass: whole GPU suite, default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r02f_pytest.log
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e large %.4e mppi %.2f ms rollout %.4e go1 %.4e cpu %.4e / 1core %.4e" % (d["value"], d["e2e"]["value"], d["large_batch"]["env_steps_per_s"],
          d["mppi"]["ms_per_plan"], d["rollout"]["env_steps_per_s_rollout"], d["step_go1"]["env_steps_per_s"], d["cpu_baseline"]["value"], d["cpu_baseline"]["one_core"]["value"]))
    print("configs3", json.dumps(d.get("configs3"), indent=1))
    print("roofline", json.dumps(d.get("roofline"), indent=1)[:1500])
except Exception as e:
    print("bench parse error", e); print(open("gpurun_out/r02f_bench.err").read()[-3000:])
PY
