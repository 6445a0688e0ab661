set -x
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r1c_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1c_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 600 gpurun_out/bench_r1c.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python tools/profile_sample.py > gpurun_out/r1c_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^(prep|lookup|fold|gemm_|pool|adagn|transpose_v|unpool|head)' -c 40 -o /tmp/r1c_eval -f python tools/profile_eval.py 1 > gpurun_out/r1c_eval_ncu.log 2>&1
ncu -i /tmp/r1c_eval.ncu-rep --page raw --csv > gpurun_out/r1c_eval_raw.csv 2>/dev/null
ls -la gpurun_out/r1c_*
