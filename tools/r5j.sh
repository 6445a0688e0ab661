#!/bin/bash
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5j_$name.json 2> gpurun_out/r5j_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5j_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('mlp_fused','gemm_unpool_out')])
PY
}
run respf1 GECCO_MLP_RESPF=1
run respf2 GECCO_MLP_RESPF=2
run respf0 GECCO_MLP_RESPF=0
run respf1b GECCO_MLP_RESPF=1
