python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -8
python tools/prof_ppo2.py 2>&1 | grep -E "^eager|^graph|kernel time|k_tanh|k_colsum|k_ppo|^\{" | cut -c1-220
