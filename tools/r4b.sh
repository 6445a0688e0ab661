#!/bin/bash
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r4b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r4b_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r4b_bench_c2.json 2> gpurun_out/r4b_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r4b_bench_c2.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], j.get('clocks')); print([(k['name'], k['ms'], k['launches']) for k in j.get('kernel_classes', [])])
PY
