#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for suf in "" "_v64"; do
  export TTM_LIB_SUFFIX=$suf
  echo "== variant '$suf'"
  timeout 600 python tools/time_objgrad.py 2>/dev/null | tail -3 | cut -c1-600
  timeout 900 python bench.py --no-inverse --no-extras --no-fit > gpurun_out/e_bench$suf.json 2>/dev/null
  python - <<PY
import json
for l in open('gpurun_out/e_bench$suf.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['e2e']['value'],1), d['parity'], {k:round(v['ms'],4) for k,v in d['roofline']['per_k'].items() if k in ('0','31','63')})
PY
done
export TTM_LIB_SUFFIX=_v64
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
