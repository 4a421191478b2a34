#!/bin/bash
# one full ncu capture of k_step at 4096 envs (after the same command ran clean). usage: tools/gpu_prof4096.sh tag
set -u
TAG=${1:-x}
mkdir -p gpurun_out
Q="python bench.py --steps 3 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu 4096"
$Q > gpurun_out/plain_4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_4096_$TAG $Q > gpurun_out/ncu_f_4096.log 2>&1
tail -c 300 gpurun_out/plain_4096.log
