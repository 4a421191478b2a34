for N in 4096 16384 65536; do
  python tools/run_go1.py $N 28
  python tools/run_go1.py $N 28 launch_fat=1 launch_lockstep=2
  python tools/run_go1.py $N 28 launch_fat=0 launch_lockstep=2
  python tools/run_go1.py $N 28 launch_fat=1 launch_lockstep=0
done
