set -x
python -m pytest tests -m gpu -q --timeout 1500 -x > gpurun_out/pytest_r2_c.log 2>&1
tail -15 gpurun_out/pytest_r2_c.log
python tools/time_objgrad.py > gpurun_out/time_objgrad_r2_c.jsonl 2> gpurun_out/time_objgrad_r2_c.err
TTM_GRAM=0 python tools/time_objgrad.py >> gpurun_out/time_objgrad_r2_c.jsonl 2>> gpurun_out/time_objgrad_r2_c.err
TTM_Q=25 python tools/time_objgrad.py >> gpurun_out/time_objgrad_r2_c.jsonl 2>> gpurun_out/time_objgrad_r2_c.err
cat gpurun_out/time_objgrad_r2_c.jsonl
tail -3 gpurun_out/time_objgrad_r2_c.err
ncu --set full --clock-control none --import-source on -k regex:objgrad_tile -s 4 -c 1 -o gpurun_out/tile_c_k0 python tools/time_objgrad.py > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:objgrad_tile -s 34 -c 1 -o gpurun_out/tile_c_k63 python tools/time_objgrad.py >> gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_c.log
python tools/time_inverse.py > gpurun_out/time_inverse_r2_c.json 2> gpurun_out/time_inverse_r2_c.err
tail -5 gpurun_out/time_inverse_r2_c.json gpurun_out/time_inverse_r2_c.err
python tools/time_inverse_fused.py > gpurun_out/time_inverse_fused_r2_c.json 2> gpurun_out/time_inverse_fused_r2_c.err
cat gpurun_out/time_inverse_fused_r2_c.json; tail -5 gpurun_out/time_inverse_fused_r2_c.err
