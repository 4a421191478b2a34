#!/bin/bash
# Multi-GPU bench on one box. Usage (through gpurun --gpus N): tools/gpu_multi.sh N
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus $N --steps 50 --warmup 12 > gpurun_out/bench_${N}gpu.log 2>&1
timeout 300 $T bench.py --gpus $N --steps 30 --warmup 12 --envs-per-gpu 65536 --rollout-envs 0 > gpurun_out/bench_${N}gpu_65536.log 2>&1
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in (f"bench_{n}gpu", f"bench_{n}gpu_65536"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, "value %.4e e2e %.4e" % (d["value"], d["e2e"]["value"]),
              {a: b for a, b in d.get("rollout", {}).items() if not isinstance(b, (dict, str))})
    except Exception as e:
        print(f, "ERR", open(f"gpurun_out/{f}.log").read()[-800:])
PY
