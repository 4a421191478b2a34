python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -8
python tools/prof_ppo2.py 2>&1 | grep -E "^eager|^graph|reduce_kernel|tanh|k_tanh|k_colsum|Self CUDA time" | cut -c1-220
