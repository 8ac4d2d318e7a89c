set -x
./tools/pipe_probe > gpurun_out/pipe_probe_r2.json; cat gpurun_out/pipe_probe_r2.json
python -m pytest tests -m gpu -q --timeout 1500 -x > gpurun_out/pytest_r2_d.log 2>&1; tail -3 gpurun_out/pytest_r2_d.log
python tools/time_objgrad.py > gpurun_out/time_objgrad_r2_d.jsonl 2> gpurun_out/time_objgrad_r2_d.err; cat gpurun_out/time_objgrad_r2_d.jsonl
python tools/time_inverse_fused.py > gpurun_out/time_inverse_fused_r2_d.json 2> gpurun_out/time_inverse_fused_r2_d.err; cat gpurun_out/time_inverse_fused_r2_d.json
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:objgrad_tile -s 64 -c 64 --csv --log-file gpurun_out/ncu_step_r2.csv python tools/objgrad_step.py > gpurun_out/ncu_step.log 2>&1
python tools/objgrad_step.py --summarise gpurun_out/ncu_step_r2.csv gpurun_out/ncu_traffic_r2.json
for cfg in "--streams 2 --bps 1" "--streams 1 --bps 2" "--streams 2 --bps 2" "--streams 3 --bps 1"; do
  python bench.py --no-inverse --no-cpu --no-fit $cfg > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_var.json'))
    print('VARIANT', d['config']['issue'][:60], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3))
except Exception as e:
    print('VARIANT failed', e); print(open('gpurun_out/bench_var.err').read()[-1500:])
PY
done
ncu --set full --clock-control none --import-source on -k regex:inverse_fused -s 2 -c 1 -o gpurun_out/invf_d python tools/time_inverse_fused.py > gpurun_out/ncu_invf.log 2>&1; tail -2 gpurun_out/ncu_invf.log
