#!/bin/bash
python -m pytest tests/test_gemm_gpu.py tests/test_bench_shape_gpu.py tests/test_denoiser_gpu.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2l_pytest.log | tail -10
for v in 2 1; do
GECCO_ANORM=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2l_bench_$v.json 2> gpurun_out/r2l_bench.err; echo "anorm=$v bench rc=$?"; tail -2 gpurun_out/r2l_bench.err; python - <<PY
import json
j=json.load(open('gpurun_out/r2l_bench_$v.json'))
print({k:j[k] for k in ('value','ms_per_step')})
print({c['name']: c['ms'] for c in j['kernel_classes']})
PY
done
