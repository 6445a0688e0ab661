#!/bin/bash
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_mlp_fused_gpu.py tests/test_bench_shape_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5g_bench.json 2> gpurun_out/r5g_bench.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r5g_bench.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', [])])
PY
