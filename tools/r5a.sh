#!/bin/bash
# round-2 session 4: A/B of the L2 prefetches in the pair GEMM (GECCO_PAIR_OPT) + GPU test suite
for opt in 0 1 2 3; do
GECCO_PAIR_OPT=$opt timeout 600 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5a_bench_opt$opt.json 2> gpurun_out/r5a_bench_opt$opt.err; echo "bench opt=$opt rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r5a_bench_opt$opt.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', [])])
PY
done
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r5a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r5a_pytest.log
