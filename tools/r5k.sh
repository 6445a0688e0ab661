#!/bin/bash
# round-2 final state: GPU tests, smoke, bench lines of every config, ncu launch list + full capture of one evaluation
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r5k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r5k_bench_c2_full.json 2> gpurun_out/r5k_bench_c2_full.err; echo "bench default rc=$?"
for c in 1 3 4 5; do timeout 600 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5k_bench_c$c.json 2> gpurun_out/r5k_bench_c$c.err; echo "bench c$c rc=$?"; done
python - <<PY
import json
for n in ('c2_full','c1','c3','c4','c5'):
    for l in open('gpurun_out/r5k_bench_%s.json' % n):
        if l.startswith('{'):
            j=json.loads(l); print(n, {k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', j['e2e']['value'], 'roofline', j['roofline'].get('kernel_class'), round(j['roofline']['frac'],3), 'whole', j['roofline'].get('whole_path_frac_of_tensor_peak'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'lib', (j.get('library_baseline') or {}))
PY
GECCO_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 330 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r5_launches_bench.log 2>&1; echo "ncu launches rc=$?"
GECCO_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_kernel|gemm_tc_kernel|pool_tc_kernel|unpool_tc_kernel|lookup_staged_kernel|head_kernel|mlp_pair_kernel|chain_kernel' -s 39 -c 39 -o /tmp/r5_eval python tools/profile_eval.py 2 > gpurun_out/r5_eval_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r5_eval_ncu.log
ncu -i /tmp/r5_eval.ncu-rep --page raw --csv > gpurun_out/r5_eval_raw.csv 2>/dev/null
ls -la gpurun_out/r5_* | head
