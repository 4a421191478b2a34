#!/bin/bash
# zero-copy A/B of step_host, then the ncu captures of the final launch defaults (65536 envs: FAT + lockstep pairs; Go1)
mkdir -p gpurun_out
timeout 300 python tools/ab_e2e.py 4096 60 2>&1 | tail -8
timeout 300 python tools/ab_e2e.py 16384 30 2>&1 | tail -8
N=65536
Q="python bench.py --steps 3 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N"
$Q > gpurun_out/plain_$N.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_${N}_r02g $Q > gpurun_out/ncu_f_$N.log 2>&1
tail -c 300 gpurun_out/plain_$N.log; echo
for N in 4096 16384; do
  python tools/run_go1.py $N 24 > gpurun_out/plain_go1_$N.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 22 -c 1 -o gpurun_out/prof_go1_${N}_r02g python tools/run_go1.py $N 24 > gpurun_out/ncu_go1_$N.log 2>&1
  cat gpurun_out/plain_go1_$N.log
done
ls -la gpurun_out/*r02g*.ncu-rep
