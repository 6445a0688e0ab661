#!/bin/bash
# final check of the committed state: whole GPU suite, smoke, default bench line, training bench line
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r5z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r5z_bench_c2_full.json 2> gpurun_out/r5z_bench_c2_full.err; echo "bench default rc=$?"
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r5z_bench_c5.json 2> gpurun_out/r5z_bench_c5.err; echo "bench c5 rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r5z_ref.json 2> gpurun_out/r5z_ref.err; echo "ref arm rc=$?"; tail -c 300 gpurun_out/r5z_ref.json
python - <<PY
import json
for n in ('c2_full','c5'):
    for l in open('gpurun_out/r5z_bench_%s.json' % n):
        if l.startswith('{'):
            j=json.loads(l); print(n, {k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', j['e2e']['value'], 'roofline', j['roofline'].get('kernel_class'), round(j['roofline']['frac'],3), 'traffic', j['roofline'].get('traffic'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'lib', (j.get('library_baseline') or {}).get('value', (j.get('library_baseline') or {}).get('bf16_autocast')), j.get('clocks'))
PY
