#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_adaptation.py -m gpu -x -q > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s.log
tail -5 gpurun_out/pytest_s.log
timeout 300 python tools/time_entf_cycle.py > gpurun_out/entf_s.json 2> gpurun_out/entf_s.err; tail -3 gpurun_out/entf_s.json
timeout 300 python tools/profile_entf_cycle.py > gpurun_out/profile_entf_s.txt 2>&1
TTM_D=256 TTM_N=200000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:basis_dense -c 1 -o gpurun_out/basis_s python tools/time_kernels.py > gpurun_out/s2.log 2>&1
echo done
