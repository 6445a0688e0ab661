#!/bin/bash
GECCO_HINT_OUT=4256 timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_bench_shape_gpu.py -q -m gpu -x 2>&1 | tail -2
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5u_$name.json 2> gpurun_out/r5u_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5u_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('mlp_fused','gemm_unpool_out','gemm_img_proj')])
PY
}
run base X=1
run late GECCO_HINT_OUT=4256
run base2 X=1
run late2 GECCO_HINT_OUT=4256
