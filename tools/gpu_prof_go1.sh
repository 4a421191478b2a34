#!/bin/bash
# ncu captures of the 3-joint-leg step kernel (Unitree Go1): 4096 envs (FAT) and 16384 envs (lean)
mkdir -p gpurun_out
for N in 4096 16384; do
  python tools/run_go1.py $N 24 > gpurun_out/plain_go1_$N.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 22 -c 1 -o gpurun_out/prof_go1_$N python tools/run_go1.py $N 24 > gpurun_out/ncu_go1_$N.log 2>&1
  cat gpurun_out/plain_go1_$N.log
done
ls -la gpurun_out/prof_go1_*
