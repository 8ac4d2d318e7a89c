#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_t.log
tail -5 gpurun_out/pytest_t.log
timeout 300 python tools/time_entf_cycle.py > gpurun_out/entf_t.json 2> gpurun_out/entf_t.err; tail -3 gpurun_out/entf_t.json
TTM_D=256 TTM_N=200000 timeout 600 python tools/time_kernels.py > gpurun_out/kernels_t_d256.json 2> gpurun_out/kernels_t.err; tail -2 gpurun_out/kernels_t_d256.json
timeout 600 python tools/time_kernels.py > gpurun_out/kernels_t_d64.json 2>> gpurun_out/kernels_t.err; tail -2 gpurun_out/kernels_t_d64.json
timeout 600 python tools/time_inverse_fused.py > gpurun_out/invf_t.json 2> gpurun_out/invf_t.err; tail -2 gpurun_out/invf_t.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:inverse_ -c 12 --csv --log-file gpurun_out/inv_traffic_t.csv python tools/time_inverse_fused.py > gpurun_out/t1.log 2>&1
echo done
