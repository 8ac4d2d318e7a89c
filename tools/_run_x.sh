#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
F="--no-cpu --no-inverse --no-extras --no-fit"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 $F > gpurun_out/x_2gpu.json 2>/dev/null
python - <<PY
import json
for l in open('gpurun_out/x_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('2gpu', d['value'], d['e2e']['value'], d['clocks'])
PY
