#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r5ac_bench_8gpu_c2.json 2> gpurun_out/r5ac_bench_8gpu_c2.err; echo "bench 8gpu rc=$?"; grep -v "OMP_NUM\|^\*" gpurun_out/r5ac_bench_8gpu_c2.err | tail -2
python - <<PY
import json
for l in open('gpurun_out/r5ac_bench_8gpu_c2.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', j['e2e']['value'], j.get('clocks'))
PY
