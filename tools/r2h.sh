#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2h_pytest.log | tail -10
python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2h_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','cuda_graph')}, 'e2e', j['e2e']['value'])
print({k:j['roofline'][k] for k in ('kernel_class','bound','achieved','frac','whole_path_frac_of_tensor_peak')})
print({c['name']: c['ms'] for c in j['kernel_classes']})
print(j['clocks'])
PY
