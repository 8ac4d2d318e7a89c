#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 1500 python bench.py > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final_1gpu.err; echo "rc=$?"
python -c "from __graft_entry__ import smoke; smoke(); print('smoke ok')" 2>&1 | tail -2
