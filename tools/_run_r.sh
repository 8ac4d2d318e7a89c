#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r.log
tail -5 gpurun_out/pytest_r.log
TTM_INV_SEARCH=1 timeout 600 python tools/time_inverse_fused.py > gpurun_out/invf_r_s1.json 2> gpurun_out/invf_r.err; tail -3 gpurun_out/invf_r_s1.json
timeout 600 python tools/time_inverse_fused.py > gpurun_out/invf_r_s2.json 2>> gpurun_out/invf_r.err; tail -3 gpurun_out/invf_r_s2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:inverse_ -c 12 --csv --log-file gpurun_out/inv_launches_r.csv python tools/time_inverse_fused.py > gpurun_out/r1.log 2>&1
TTM_NS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:inverse_fused --launch-skip 5 -c 1 -o gpurun_out/invwalk_r python tools/time_inverse_fused.py > gpurun_out/r3.log 2>&1
echo done
