#!/bin/bash
# final round-2 ncu captures of the shipped step kernel: 4096 envs (FAT instantiation) and 65536 envs (lean, lockstep pairs),
# and the launch list of the default 4096-env bench command. Each capture runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
for N in 4096 65536; do
  Q="python bench.py --steps 3 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu $N"
  $Q > gpurun_out/plain_$N.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_${N}_r02f $Q > gpurun_out/ncu_f_$N.log 2>&1
  tail -c 200 gpurun_out/plain_$N.log; echo
done
Q4="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0"
$Q4 > gpurun_out/plain_4096b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4096_r02f.csv $Q4 > gpurun_out/ncu_l_4096.log 2>&1
wc -l gpurun_out/launches_4096_r02f.csv; ls -la gpurun_out/*.ncu-rep
