#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
N=${NG:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu > gpurun_out/bench_final_${N}gpu.json 2> gpurun_out/bench_final_${N}gpu.err; echo "rc=$?"
tail -c 300 gpurun_out/bench_final_${N}gpu.json
