set -x
python -m pytest tests -m gpu -q --timeout 1500 -x > gpurun_out/pytest_r2_g.log 2>&1; tail -3 gpurun_out/pytest_r2_g.log
TTM_N=1000 python tools/profile_entf_cycle.py > gpurun_out/profile_entf_r2.txt 2>&1; tail -60 gpurun_out/profile_entf_r2.txt
python bench.py --no-inverse --no-fit --steps 3 > gpurun_out/bench_r2_g.json 2> gpurun_out/bench_r2_g.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_g.json'))
print(json.dumps(d.get('other_configs'), indent=1)[:6000])
PY
tail -3 gpurun_out/bench_r2_g.err
TTM_D=64 python tools/time_kernels.py > gpurun_out/kernels_r2_d64.json 2> gpurun_out/kernels_r2_d64.err; cat gpurun_out/kernels_r2_d64.json; tail -3 gpurun_out/kernels_r2_d64.err
