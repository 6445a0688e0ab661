#!/bin/bash
for mp in 1 0; do
GECCO_MLP_PAIR=$mp timeout 600 python bench.py --steps 3 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r3e_bench_mp$mp.json 2> gpurun_out/r3e_bench_mp$mp.err; echo "bench mlp_pair=$mp rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r3e_bench_mp$mp.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], j.get('clocks')); print([(k['name'], k['ms'], k['launches']) for k in j.get('kernel_classes', [])])
PY
done
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3e_pytest.log
