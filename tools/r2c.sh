#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_pytest.log | tail -30
python -m pytest tests/test_bench_shape_gpu.py -m gpu -q -s 2>&1 | grep -E "rel rms|loss|sampler|ConvNeXt|error" > gpurun_out/r2c_bench_shape_numbers.log; cat gpurun_out/r2c_bench_shape_numbers.log | head -60
python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2c_bench.json'))
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','cuda_graph')})
    print('e2e',j['e2e']['value'],'roofline',{k:j['roofline'][k] for k in ('kernel_class','bound','achieved','frac','whole_path_frac_of_tensor_peak')})
    for c in j['kernel_classes']: print(c)
    print(j.get('cpu_baseline')); print(j.get('library_baseline'))
except Exception as e: print('bench parse failed',e)
PY
tail -5 gpurun_out/r2c_bench.err
GECCO_ANORM=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2c_bench_fold.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r2c_bench_fold.json')); print('fold path:', j['value'], j['ms_per_step'])"
