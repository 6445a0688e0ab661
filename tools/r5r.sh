#!/bin/bash
GECCO_TC_REV=1 timeout 300 python -m pytest tests/test_bench_shape_gpu.py tests/test_denoiser_gpu.py -q -m gpu -x 2>&1 | tail -2
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5r_$name.json 2> gpurun_out/r5r_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5r_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('lookup','fold_group_norm','gemm_img_proj','gemm_kv_q','head_edm_step')])
PY
}
run base X=1
run tcrev GECCO_TC_REV=1
run base2 X=1
run tcrev2 GECCO_TC_REV=1
