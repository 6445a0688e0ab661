#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2o_pytest.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2o_bench.err; python - <<PY
import json
j=json.load(open('gpurun_out/r2o_bench.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','cuda_graph')}, 'e2e', j['e2e']['value'])
print({k:j['roofline'][k] for k in ('kernel_class','bound','achieved','frac','whole_path_frac_of_tensor_peak')}, j['roofline']['all_tcgen05_gemms'])
print({c['name']: c['ms'] for c in j['kernel_classes']})
print(j['clocks'], j.get('cpu_baseline',{}).get('value'), j.get('library_baseline'))
PY
