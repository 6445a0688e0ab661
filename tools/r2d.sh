#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_bench_shape_gpu.py -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2d_pytest.log | tail -10
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2d_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step','cuda_graph')})
for c in j['kernel_classes']: print(c['name'], c['ms'], c['tflops'], c['gbs'])
PY
