timeout 300 bash tools/gpu_tests_isolated.sh tests/test_ops_gpu.py -k "pool_attention" 2>&1 | tail -12
timeout 120 python tools/pool_time.py 2>&1 | tail -5
timeout 400 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -c 1500 gpurun_out/bench_h.json
