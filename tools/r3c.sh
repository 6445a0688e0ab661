#!/bin/bash
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3c_pytest.log
for c in 2 1; do
timeout 600 python bench.py --steps 3 --warmup 3 --config $c --no-cpu-baseline > gpurun_out/r3c_bench_c$c.json 2> gpurun_out/r3c_bench_c$c.err; echo "bench config $c rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r3c_bench_c$c.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value']); print([(k['name'], k['ms']) for k in j.get('kernel_classes', [])])
PY
done
