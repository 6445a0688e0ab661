#!/bin/bash
# upsample through graphs + config 4 bench, then the round-2 ncu evidence
python -m pytest tests/test_denoiser_gpu.py -m gpu -q -k "cond_uvl or uncond" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2j_pytest.log
python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err; echo "config 4 bench rc=$?"; tail -3 gpurun_out/r2j_bench_c4.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2j_bench_c4.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','cuda_graph')}, 'e2e', j['e2e']['value'], 'whole', j['roofline']['whole_path_frac_of_tensor_peak'])
PY
# launch list of the bench command (eager launches so that every kernel is a separate ncu record)
GECCO_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu launches rc=$?"
# full capture of one evaluation's tensor-core / gather / head kernels
GECCO_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_kernel|gemm_tc_kernel|pool_tc_kernel|unpool_tc_kernel|lookup_staged_kernel|head_kernel' -s 38 -c 38 -o gpurun_out/r2_eval python tools/profile_eval.py 2 > gpurun_out/r2_eval_ncu.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2_eval_ncu.log
ls -la gpurun_out/r2_eval.ncu-rep gpurun_out/r2_launches.csv
