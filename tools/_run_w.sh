#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
F="--no-cpu --no-inverse --no-extras --no-fit"
python bench.py $F > gpurun_out/w_plain.json 2>/dev/null
OMP_NUM_THREADS=1 python bench.py $F > gpurun_out/w_omp1.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 1 $F > gpurun_out/w_torchrun1.json 2>/dev/null
for f in plain omp1 torchrun1; do python - <<PY
import json
for l in open('gpurun_out/w_$f.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', d['value'], d['e2e']['value'])
PY
done
