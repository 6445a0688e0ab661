#!/bin/bash
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5s_$name.json 2> gpurun_out/r5s_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5s_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('mlp_fused','gemm_unpool_out','gemm_kv_q')])
PY
}
run base X=1
run h517 GECCO_HINT_MLP=517
run h2565 GECCO_HINT_MLP=2565
run h1541 GECCO_HINT_MLP=1541
run base2 X=1
