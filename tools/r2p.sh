#!/bin/bash
# 2-GPU weak scaling with the final all-gather inside the timed region: config 2 (headline) and config 3 (BASELINE.json's sharded config)
for c in 2 3; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --config $c --no-cpu-baseline > gpurun_out/r2p_bench_2gpu_c$c.json 2> gpurun_out/r2p_bench_2gpu_c$c.err; echo "config $c 2-GPU rc=$?"; tail -2 gpurun_out/r2p_bench_2gpu_c$c.err
python - <<PY
import json
for l in open('gpurun_out/r2p_bench_2gpu_c$c.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j[k] for k in ('value','n_gpus','ms_per_step','host_enqueue_ms_per_step','cuda_graph')}, 'e2e', j['e2e']['value'], j['config']['parallelism'])
PY
done
python bench.py --steps 3 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/r2p_bench_1gpu_c2.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r2p_bench_1gpu_c2.json')); print('1 GPU same box:', j['value'], j['ms_per_step'])"
