timeout 300 bash tools/gpu_tests_isolated.sh tests/test_ops_gpu.py -k "pool_attention or lookup" 2>&1 | tail -30
timeout 120 python tools/lookup_time.py 2>&1 | tail -9
timeout 120 python tools/pool_time.py 2>&1 | tail -5
