#!/bin/bash
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5o_$name.json 2> gpurun_out/r5o_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5o_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('mlp_fused','gemm_unpool_out','gemm_kv_q','unpool_attention')])
PY
}
run base X=1
run out160_mlp5 GECCO_HINT_OUT=160 GECCO_HINT_MLP=5
run mlp5 GECCO_HINT_MLP=5
run mlp4 GECCO_HINT_MLP=4
run mlp1 GECCO_HINT_MLP=1
run out32 GECCO_HINT_OUT=32
run out128 GECCO_HINT_OUT=128
run un128 GECCO_HINT_UNPOOL=128
run base2 X=1
