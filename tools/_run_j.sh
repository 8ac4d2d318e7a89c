set -x
python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/pytest_r2_j.log 2>&1; tail -12 gpurun_out/pytest_r2_j.log
python bench.py > gpurun_out/bench_r2_j.json 2> gpurun_out/bench_r2_j.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_j.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'fit',d.get('fit'))
print('parity', d.get('parity'))
i=d['inverse_map']; print('inv e2e',i['value'],'device',i['device']['samples_per_s'],'ctor',i['ctor_s'],'opt',i['optimize_s'])
print(json.dumps(d['other_configs']['C3_example06_entf_cycle']))
print(json.dumps(d['other_configs']['C2_example05_densities'])[:600])
PY
tail -3 gpurun_out/bench_r2_j.err
