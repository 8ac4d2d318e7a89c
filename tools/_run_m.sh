set -x
python -m pytest tests -m gpu -q --timeout 1500 -s > gpurun_out/pytest_r2_m.log 2>&1; grep -v "^$" gpurun_out/pytest_r2_m.log | tail -14
python bench.py > gpurun_out/bench_r2_m.json 2> gpurun_out/bench_r2_m.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_m.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'fit',d.get('fit'))
print('parity', d.get('parity'))
i=d['inverse_map']; print('inv e2e',i['value'],'pageable',i['table_pageable']['samples_per_s'],'device',i['device']['samples_per_s'],'ctor',i['ctor_s'],'opt',i['optimize_s'])
PY
tail -3 gpurun_out/bench_r2_m.err
