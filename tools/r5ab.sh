#!/bin/bash
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5ab_$name.json 2> gpurun_out/r5ab_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5ab_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz'))
PY
}
run base X=1
run smallpdl GECCO_SMALL_PDL=1
run tcrev GECCO_TC_REV=1
run base2 X=1
