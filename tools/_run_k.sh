set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err
tail -c 1500 gpurun_out/bench_r2_8gpu.json; tail -5 gpurun_out/bench_r2_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r2_4gpu.json 2> gpurun_out/bench_r2_4gpu.err
tail -c 600 gpurun_out/bench_r2_4gpu.json
