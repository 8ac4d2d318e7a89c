set -x
python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/pytest_r2_h.log 2>&1; tail -25 gpurun_out/pytest_r2_h.log
TTM_D=64 python tools/time_kernels.py > gpurun_out/kernels_r2_d64.json 2> gpurun_out/kernels_r2_d64.err; cat gpurun_out/kernels_r2_d64.json; tail -3 gpurun_out/kernels_r2_d64.err
