#!/bin/bash
# A/B: lookup channel slices (2 CTAs per SM at S = 4), pool key splits
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5d_$name.json 2> gpurun_out/r5d_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5d_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('lookup','pool_attention','inducer_chain')])
PY
}
run base X=1
run slices4 GECCO_LOOKUP_SLICES=4
run splits2 GECCO_POOL_TC_SPLITS=2
run splits3 GECCO_POOL_TC_SPLITS=3
