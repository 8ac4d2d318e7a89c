set -x
nvidia-smi -L
export TTM_MULTI_GPU_LOG=$PWD/gpurun_out/multi_gpu_check_r2.txt
python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/pytest_r2_f.log 2>&1; tail -8 gpurun_out/pytest_r2_f.log
cat gpurun_out/multi_gpu_check_r2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err
tail -c 3000 gpurun_out/bench_r2_2gpu.json; tail -5 gpurun_out/bench_r2_2gpu.err
