#!/bin/bash
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5h_$name.json 2> gpurun_out/r5h_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5h_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'].startswith('gemm')])
PY
}
run kbps1 GECCO_PAIR_KBPS=1
run kbps3 GECCO_PAIR_KBPS=3
run base X=1
