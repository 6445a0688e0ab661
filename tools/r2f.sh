#!/bin/bash
python -m pytest tests/test_gemm_gpu.py tests/test_bench_shape_gpu.py tests/test_denoiser_gpu.py -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2f_pytest.log | tail -10
for v in "1 1" "1 0" "0 1" "0 0"; do set -- $v
GECCO_ANORM=$1 GECCO_FAST_EPILOGUE=$2 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2f_bench_$1$2.json 2> gpurun_out/r2f_bench.err; echo "anorm=$1 fast=$2 bench rc=$?"; python - <<PY
import json
j=json.load(open('gpurun_out/r2f_bench_$1$2.json'))
print({k:j[k] for k in ('value','ms_per_step')})
print({c['name']: c['ms'] for c in j['kernel_classes'] if c['name'].startswith('gemm') or c['name'].startswith('fold_a')})
PY
done
