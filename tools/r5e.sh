#!/bin/bash
timeout 400 python -m pytest tests/test_training_gpu.py tests/test_samplers_gpu.py -x -q -s > gpurun_out/r5e_train.log 2>&1; echo "train tests rc=$?"; grep -E "GroupAffineNorm|GaussAct|losses|passed|failed|Error|grad" gpurun_out/r5e_train.log | tail -20
for fused in 1; do
GECCO_TRAIN_FUSED=$fused timeout 400 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r5e_bench_c5_f$fused.json 2> gpurun_out/r5e_bench_c5_f$fused.err; echo "bench c5 fused=$fused rc=$?"; tail -3 gpurun_out/r5e_bench_c5_f$fused.err
python - <<PY
import json
for l in open('gpurun_out/r5e_bench_c5_f$fused.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches','loss','cuda_graph')}, 'e2e', j['e2e']['value'], j['roofline']['frac'], j['library_baseline'])
PY
done
