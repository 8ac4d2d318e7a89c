#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:inverse_ --csv --log-file gpurun_out/inv_launches_p.csv python tools/time_inverse_fused.py > gpurun_out/p1.log 2>&1
TTM_NS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:inverse_rect -c 1 -o gpurun_out/invrect_p python tools/time_inverse_fused.py > gpurun_out/p2.log 2>&1
TTM_NS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:inverse_fused --launch-skip 5 -c 1 -o gpurun_out/invwalk_p python tools/time_inverse_fused.py > gpurun_out/p3.log 2>&1
echo done
