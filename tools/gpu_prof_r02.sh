#!/bin/bash
# round-2 ncu captures: k_step at 65536 envs, k_mlp at 16384 rows, launch list of the default 4096-env command
set -u
mkdir -p gpurun_out
Q="python bench.py --steps 3 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0 --envs-per-gpu 65536"
$Q > gpurun_out/plain_65536.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 13 -c 1 -o gpurun_out/prof_65536_r02a $Q > gpurun_out/ncu_f_65536.log 2>&1
tail -c 200 gpurun_out/plain_65536.log; echo
python tools/run_mlp.py 16384 > gpurun_out/plain_mlp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mlp -s 8 -c 1 -o gpurun_out/prof_mlp_r02a python tools/run_mlp.py 16384 > gpurun_out/ncu_f_mlp.log 2>&1
cat gpurun_out/plain_mlp.log
Q4="python bench.py --steps 20 --warmup 12 --no-cpu-baseline --large-batch 0 --rollout-envs 0 --train-envs 0 --mppi 0 --go1 0"
$Q4 > gpurun_out/plain_4096b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4096_r02a.csv $Q4 > gpurun_out/ncu_l_4096.log 2>&1
tail -c 200 gpurun_out/plain_4096b.log; echo; wc -l gpurun_out/launches_4096_r02a.csv
