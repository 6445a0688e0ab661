#!/bin/bash
GECCO_TRAIN_COND_AUTOCAST=1 timeout 300 python -m pytest tests/test_training_gpu.py -x -q -s -k "loss_and_gradients" 2>&1 | grep "grads_\|passed\|failed\|assert" | head -12
for ac in 1 0; do
GECCO_TRAIN_COND_AUTOCAST=$ac timeout 400 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5y_c5_ac$ac.json 2> gpurun_out/r5y_c5_ac$ac.err; echo "bench c5 autocast=$ac rc=$?"; tail -2 gpurun_out/r5y_c5_ac$ac.err
python - <<PY
import json
for l in open('gpurun_out/r5y_c5_ac$ac.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step','loss')})
PY
done
