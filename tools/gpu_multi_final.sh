#!/bin/bash
# the driver's N-GPU bench command on one box, and its headline / configs[3] numbers. Usage (gpurun --gpus N): tools/gpu_multi_final.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 12 > gpurun_out/bench_${N}gpu_final.log 2> gpurun_out/bench_${N}gpu_final.err
echo "exit $?"; python tools/show_multi.py gpurun_out/bench_${N}gpu_final.log || tail -c 1500 gpurun_out/bench_${N}gpu_final.err
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_{sys.argv[1]}gpu_final.log").read().strip().splitlines()[-1])
print("each", d["configs3"]["ms_each_iteration"], "host", d["configs3"].get("ms_host"))
PY
