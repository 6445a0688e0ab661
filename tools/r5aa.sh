#!/bin/bash
for gm in 1 0; do
GECCO_TRAIN_GRAPH=$gm timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$gm bench.py --gpus 2 --config 5 --steps 5 --warmup 3 > gpurun_out/r5aa_c5_2gpu_g$gm.json 2> gpurun_out/r5aa_c5_2gpu_g$gm.err; echo "bench c5 2gpu graph=$gm rc=$?"; grep -v "OMP_NUM\|^\*" gpurun_out/r5aa_c5_2gpu_g$gm.err | tail -2
python - <<PY
import json
for l in open('gpurun_out/r5aa_c5_2gpu_g$gm.json'):
    if l.startswith('{'):
        j=json.loads(l); print({k:j.get(k) for k in ('value','n_gpus','ms_per_step','gpu_launches','loss','cuda_graph')}, 'e2e', j['e2e']['value'], j['config']['parallelism'])
PY
done
