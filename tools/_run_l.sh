set -x
python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/pytest_r2_l.log 2>&1; tail -12 gpurun_out/pytest_r2_l.log
TTM_D=64 python tools/time_kernels.py > gpurun_out/kernels_r2_d64.json 2> gpurun_out/kernels_r2_d64.err; python -c "
import json; d=json.load(open('gpurun_out/kernels_r2_d64.json')); print(d['basis'], d['gram'])"
python - <<'PY'
import sys, time, json
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from cases import synthetic_samples, c4_terms
from transport_map import transport_map
D,n=64,1_000_000
mon,non=c4_terms(D)
tm=transport_map(X=synthetic_samples(n,D,seed=0),monotone=mon,nonmonotone=non,monotonicity='integrated rectifier',quadrature_input={'order':100},verbose=False)
for thr in (1,2,4):
    tm.fit_threads=thr
    for k in range(D):
        tm.coeffs_mon[k]*=0; tm.coeffs_nonmon[k]*=0
    tm._fg_cache={}
    t=time.perf_counter(); tm.optimize(); torch.cuda.synchronize(); dt=time.perf_counter()-t
    info=tm._fit_info
    print('threads',thr,'fit_s',round(dt,3),'evals',sum(info[k]['nfev'] for k in range(D)),'sum comp s',round(sum(info[k]['seconds'] for k in range(D)),3))
    print('   per k (k, nfev, nit, ms):',[(k,info[k]['nfev'],info[k]['nit'],round(info[k]['seconds']*1e3,1)) for k in (0,1,15,31,47,63)])
PY
