#!/bin/bash
timeout 400 python -m pytest tests/test_denoiser_gpu.py tests/test_engine_state_gpu.py tests/test_bench_shape_gpu.py -q -m gpu -x 2>&1 | tail -3
for kv in 1 0; do
GECCO_UPSAMPLE_KV=$kv timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5q_c4_kv$kv.json 2> gpurun_out/r5q_c4_kv$kv.err; echo "bench c4 kv=$kv rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r5q_c4_kv$kv.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, [(k['name'], k['launches'], k['ms']) for k in j['kernel_classes'] if k['name'] in ('inducer_chain','unpool_attention')])
PY
done
