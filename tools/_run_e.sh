set -x
./tools/pipe_probe > gpurun_out/pipe_probe_r2.json; cat gpurun_out/pipe_probe_r2.json
python -m pytest tests -m gpu -q --timeout 1500 -x > gpurun_out/pytest_r2_e.log 2>&1; tail -3 gpurun_out/pytest_r2_e.log
python tools/time_inverse_fused.py > gpurun_out/time_inverse_fused_r2_e.json 2> gpurun_out/time_inverse_fused_r2_e.err; cat gpurun_out/time_inverse_fused_r2_e.json; tail -3 gpurun_out/time_inverse_fused_r2_e.err
python bench.py > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err; cat gpurun_out/bench_r2_e.json; tail -5 gpurun_out/bench_r2_e.err
