#!/bin/bash
# run O: split inverse + lockstep L-BFGS-B
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_o.log
tail -5 gpurun_out/pytest_o.log
timeout 600 python tools/time_inverse_fused.py > gpurun_out/invf_o.json 2> gpurun_out/invf_o.err; tail -3 gpurun_out/invf_o.json
timeout 300 python tools/time_entf_cycle.py > gpurun_out/entf_o.json 2> gpurun_out/entf_o.err; tail -3 gpurun_out/entf_o.json
TTM_HOST_OPT=scipy timeout 300 python tools/time_entf_cycle.py > gpurun_out/entf_o_scipy.json 2>&1; tail -3 gpurun_out/entf_o_scipy.json
timeout 300 python tools/profile_entf_cycle.py > gpurun_out/profile_entf_o.txt 2>&1
TTM_NS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:inverse_ -c 4 -o gpurun_out/invsplit_o python tools/time_inverse_fused.py > gpurun_out/ncu_o.log 2>&1
echo done
