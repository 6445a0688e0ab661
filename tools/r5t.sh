#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r5t_bench_2gpu_c2.json 2> gpurun_out/r5t_bench_2gpu_c2.err; echo "bench 2gpu rc=$?"; tail -2 gpurun_out/r5t_bench_2gpu_c2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r5t_ref_2gpu.json 2> gpurun_out/r5t_ref_2gpu.err; echo "ref arm 2gpu rc=$?"; tail -c 600 gpurun_out/r5t_ref_2gpu.json
python - <<PY
import json
for l in open('gpurun_out/r5t_bench_2gpu_c2.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','n_gpus','ms_per_step','scaling')}, 'e2e', j['e2e']['value'], j['config'].get('parallelism'))
PY
