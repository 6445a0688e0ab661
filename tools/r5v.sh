#!/bin/bash
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_chain_gpu.py tests/test_bench_shape_gpu.py tests/test_denoiser_gpu.py -q -m gpu -x 2>&1 | tail -3
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5v_$name.json 2> gpurun_out/r5v_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5v_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('pool_attention','inducer_chain','gemm_kv_q','unpool_attention')])
PY
}
run uneven1 GECCO_POOL_UNEVEN=1
run uneven0 GECCO_POOL_UNEVEN=0
run uneven1b GECCO_POOL_UNEVEN=1
run uneven0b GECCO_POOL_UNEVEN=0
