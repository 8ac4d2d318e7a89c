set -x
python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/pytest_r2_i.log 2>&1; tail -12 gpurun_out/pytest_r2_i.log
python tools/time_objgrad.py > gpurun_out/time_objgrad_r2_i.jsonl 2> gpurun_out/time_objgrad_r2_i.err; cat gpurun_out/time_objgrad_r2_i.jsonl; tail -2 gpurun_out/time_objgrad_r2_i.err
python tools/time_inverse.py > gpurun_out/time_inverse_r2_i.json 2> gpurun_out/time_inverse_r2_i.err; cat gpurun_out/time_inverse_r2_i.json; tail -2 gpurun_out/time_inverse_r2_i.err
python bench.py --no-inverse --no-cpu --no-extras > gpurun_out/bench_r2_i.json 2> gpurun_out/bench_r2_i.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_i.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'fit',d.get('fit'))
print({k:v['ms'] for k,v in d['roofline']['per_k'].items()})
PY
tail -3 gpurun_out/bench_r2_i.err
