#!/bin/bash
GECCO_HINT_OUT=165 GECCO_HINT_UNPOOL=129 GECCO_HINT_MLP=149 GECCO_HINT_KVQ=129 timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_bench_shape_gpu.py tests/test_mlp_fused_gpu.py -q -m gpu -x 2>&1 | tail -3
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r5n_$name.json 2> gpurun_out/r5n_$name.err; echo "bench $name $@ rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r5n_$name.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','ms_per_step')}, j.get('clocks',{}).get('sm_mhz')); print([(k['name'], k['ms']) for k in j.get('kernel_classes', []) if k['name'] in ('mlp_fused','gemm_unpool_out','gemm_kv_q','pool_attention','unpool_attention')])
PY
}
run base X=1
run out160 GECCO_HINT_OUT=160
run out165 GECCO_HINT_OUT=165
run out165_un129 GECCO_HINT_OUT=165 GECCO_HINT_UNPOOL=129
run plus_mlp5 GECCO_HINT_OUT=165 GECCO_HINT_UNPOOL=129 GECCO_HINT_MLP=5
run plus_mlp149 GECCO_HINT_OUT=165 GECCO_HINT_UNPOOL=129 GECCO_HINT_MLP=149
run plus_kvq GECCO_HINT_OUT=165 GECCO_HINT_UNPOOL=129 GECCO_HINT_MLP=149 GECCO_HINT_KVQ=129
run base2 X=1
