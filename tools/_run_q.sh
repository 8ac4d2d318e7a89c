#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -5 gpurun_out/pytest_q.log
timeout 600 python tools/time_inverse_fused.py > gpurun_out/invf_q.json 2> gpurun_out/invf_q.err; tail -3 gpurun_out/invf_q.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:inverse_ -c 12 --csv --log-file gpurun_out/inv_launches_q.csv python tools/time_inverse_fused.py > gpurun_out/q1.log 2>&1
TTM_NS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:inverse_rect -c 1 -o gpurun_out/invrect_q python tools/time_inverse_fused.py > gpurun_out/q2.log 2>&1
TTM_NS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:inverse_fused --launch-skip 5 -c 1 -o gpurun_out/invwalk_q python tools/time_inverse_fused.py > gpurun_out/q3.log 2>&1
echo done
