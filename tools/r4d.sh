#!/bin/bash
timeout 900 python -m pytest tests/test_training_gpu.py -x -q -s > gpurun_out/r4d_train.log 2>&1; echo "train tests rc=$?"; grep -E "losses|passed|failed|Error" gpurun_out/r4d_train.log | tail
for gm in 1 0; do
GECCO_TRAIN_GRAPH=$gm timeout 900 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r4d_bench_c5_g$gm.json 2> gpurun_out/r4d_bench_c5_g$gm.err; echo "bench c5 graph=$gm rc=$?"; tail -3 gpurun_out/r4d_bench_c5_g$gm.err
python - <<PY
import json
for l in open('gpurun_out/r4d_bench_c5_g$gm.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches','loss','peak_memory_gb','cuda_graph')}, 'e2e', j['e2e']['value'], j['roofline']['frac'], j['library_baseline'])
PY
done
