#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; echo "rc=$?"
tail -c 3000 gpurun_out/bench_u.json
tail -5 gpurun_out/bench_u.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_u_ref.json 2> gpurun_out/bench_u_ref.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_u_ref.json
