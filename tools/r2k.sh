#!/bin/bash
# round-2 ncu evidence (reports stay on the box; CSV exports come back) + hybrid lookup check
python -m pytest tests/test_ops_gpu.py -m gpu -q -k lookup > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2k_pytest.log
python bench.py --config 3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err; echo "config 3 bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2k_bench_c3.json'))
print({k:j[k] for k in ('value','ms_per_step')}, j['roofline']['lookup_hbm'])
PY
GECCO_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu launches rc=$?"
GECCO_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_kernel|gemm_tc_kernel|pool_tc_kernel|unpool_tc_kernel|lookup_staged_kernel|head_kernel' -s 38 -c 38 -o /tmp/r2_eval python tools/profile_eval.py 2 > gpurun_out/r2_eval_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2_eval_ncu.log
ncu -i /tmp/r2_eval.ncu-rep --page raw --csv > gpurun_out/r2_eval_raw.csv 2>/dev/null
ncu -i /tmp/r2_eval.ncu-rep --page source --csv -k regex:gemm_pair_kernel -c 1 > gpurun_out/r2_kvq_source.csv 2>/dev/null
ls -la gpurun_out/
