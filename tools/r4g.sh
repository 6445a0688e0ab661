#!/bin/bash
# round-2 (final state) ncu evidence: launch list of the bench command + full capture of one evaluation
GECCO_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 330 --csv --log-file gpurun_out/r4_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_launches_bench.log 2>&1; echo "ncu launches rc=$?"
GECCO_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_kernel|gemm_tc_kernel|pool_tc_kernel|unpool_tc_kernel|lookup_staged_kernel|head_kernel|mlp_pair_kernel|chain_kernel' -s 39 -c 39 -o /tmp/r4_eval python tools/profile_eval.py 2 > gpurun_out/r4_eval_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r4_eval_ncu.log
ncu -i /tmp/r4_eval.ncu-rep --page raw --csv > gpurun_out/r4_eval_raw.csv 2>/dev/null
ncu -i /tmp/r4_eval.ncu-rep --page source --csv -k regex:mlp_pair_kernel -c 1 > gpurun_out/r4_mlp_pair_source.csv 2>/dev/null
ls -la gpurun_out/r4_*
