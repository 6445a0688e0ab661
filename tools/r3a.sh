#!/bin/bash
# round 3 (session 2 of round 2): inducer chain kernel: op tests, whole GPU suite, bench A/B
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -m gpu > gpurun_out/r3a_chain.log 2>&1; echo "chain tests rc=$?"; tail -15 gpurun_out/r3a_chain.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3a_pytest.log
for ch in 1 0; do
GECCO_CHAIN=$ch timeout 600 python bench.py --steps 3 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r3a_bench_chain$ch.json 2> gpurun_out/r3a_bench_chain$ch.err; echo "bench chain=$ch rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r3a_bench_chain$ch.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value']); print(j.get('kernel_classes'))
PY
done
GECCO_CHAIN=1 timeout 600 python bench.py --steps 3 --warmup 3 --config 1 --no-cpu-baseline > gpurun_out/r3a_bench_c1.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r3a_bench_c1.json')); print('config 1:', j['value'], j['ms_per_step'])"
