#!/bin/bash
python -m pytest tests/test_samplers_gpu.py tests/test_denoiser_gpu.py tests/test_modules_gpu.py tests/test_engine_state_gpu.py -m gpu -q -s > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error|rel rms" gpurun_out/r2i_pytest.log | tail -30
for c in 4 3 1; do
python bench.py --config $c --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2i_bench_c$c.json 2> gpurun_out/r2i_bench_c$c.err; echo "config $c bench rc=$?"; tail -3 gpurun_out/r2i_bench_c$c.err; python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2i_bench_c$c.json'))
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','cuda_graph')}, 'e2e', j['e2e']['value'])
    print({k:j['roofline'][k] for k in ('kernel_class','bound','achieved','frac','whole_path_frac_of_tensor_peak')}, j['roofline']['lookup_hbm'])
    print({c['name']: c['ms'] for c in j['kernel_classes']})
except Exception as e: print('parse failed', e)
PY
done
