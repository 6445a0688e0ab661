#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2b_smoke.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cat gpurun_out/r2b_bench.json | cut -c1-900; tail -5 gpurun_out/r2b_bench.err
