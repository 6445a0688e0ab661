#!/bin/bash
python -m pytest tests/test_gemm_gpu.py tests/test_bench_shape_gpu.py -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2n_pytest.log | tail -6
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2n_bench.err; python - <<PY
import json
j=json.load(open('gpurun_out/r2n_bench.json'))
print({k:j[k] for k in ('value','ms_per_step')})
print({c['name']: c['ms'] for c in j['kernel_classes']})
PY
GECCO_DEBUG_COUNTERS=1 python -m gecco_b200.build > /dev/null 2>&1
CASES=kvq_anorm,mlp_up_anorm python tools/gemm_cycles.py 2>&1 | grep -E "==|leader" | cut -c1-700
