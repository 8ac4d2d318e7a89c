#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
N=${NG:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu > gpurun_out/bench_v_${N}gpu.json 2> gpurun_out/bench_v_${N}gpu.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_v_${N}gpu.json
tail -3 gpurun_out/bench_v_${N}gpu.err
TTM_MULTI_GPU_LOG=gpurun_out/multi_gpu_check_v_${N}.txt timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py > gpurun_out/check_v_${N}.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/check_v_${N}.log
