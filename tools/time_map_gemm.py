"""Forward map of the C5 map (D = 256 separable, order-3 Hermite-function nonmonotone terms): GEMM form
(ttm_map_rect + ttm_sep_eval_base) against the per-component kernels, through map() with host arrays and
device-resident (CUDA events around the kernels only)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import synthetic_samples, c5_terms, headline_sep_coeffs   # noqa: E402
from transport_map import transport_map                  # noqa: E402
from ttt_b200 import binding as B                        # noqa: E402

D, n = int(os.environ.get('TTM_D', 256)), int(os.environ.get('TTM_NS', 400_000))
mon, non = c5_terms(D)
tm = transport_map(X=synthetic_samples(4000, D, seed=0), monotone=mon, nonmonotone=non,
                   monotonicity='separable monotonicity', verbose=False)
cm, cn = headline_sep_coeffs(mon, non)
for k in range(D):
    tm.coeffs_mon[k], tm.coeffs_nonmon[k] = cm[k].copy(), cn[k].copy()
X = synthetic_samples(n, D, seed=3)
out = {'D': D, 'n': n}
res = {}
for mode in ('1', '0'):
    os.environ['TTM_MAP_GEMM'] = mode
    tm._inv_pack_cache.pop('map_gemm', None)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        Z = tm.map(X)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    res[mode] = Z
    out['e2e_s_gemm' if mode == '1' else 'e2e_s_per_component'] = min(ts)
    out['e2e_all_' + mode] = [round(t, 4) for t in ts]
out['max_rel_diff'] = float(np.max(np.abs(res['1'] - res['0'])) / np.max(np.abs(res['0'])))
# device-resident: kernels only
Xt = tm._to_colmajor(X, tm._mean_d, tm._std_d)
Zt = tm._empty(D, n)
gm_t = []
os.environ['TTM_MAP_GEMM'] = '1'
tm._inv_pack_cache.pop('map_gemm', None)
gm = tm._map_gemm_static()
cat = np.concatenate([np.asarray(tm.coeffs_nonmon[k], dtype=np.float64) for k in range(D)])
R = np.zeros(gm['r_size'])
R[gm['rdst']] = (cat[gm['src']] * gm['sc'])[gm['rkeep']]
Rd = tm._upload(R)
base = tm._empty(D, (n + 1) // 2 * 2)
st = tm._stream()
for k in range(D):
    tm._set_coeffs(k, tm.coeffs_nonmon[k], tm.coeffs_mon[k])
for rep in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    B.check(tm._lib.ttm_map_rect(tm._ctx, B.c_void_p(Xt.data_ptr()), Xt.shape[1], n, D, gm['rows'], gm['ns'],
                                 tm.skip_dimensions, B.c_void_p(Rd.data_ptr()), B.c_void_p(base.data_ptr()), base.shape[1], st))
    e1.record()
    for k in range(D):
        B.check(tm._lib.ttm_sep_eval_base(tm._plans[k], B.c_void_p(Xt.data_ptr()), Xt.shape[1], n,
                                          B.c_void_p(base[k].data_ptr()), 0.0, B.c_void_p(Zt[k].data_ptr()), st))
    e2.record()
    torch.cuda.synchronize()
    gm_t.append((e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3))
out['device_gemm_s'], out['device_mon_s'] = min(t[0] for t in gm_t), min(t[1] for t in gm_t)
flops = 2.0 * 3 * sum(range(D)) * n
out['gemm_useful_tflops'] = flops / out['device_gemm_s'] / 1e12
pc = []
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(D):
        tm._s_device(k, Xt, n, Zt[k])
    e1.record()
    torch.cuda.synchronize()
    pc.append(e0.elapsed_time(e1) * 1e-3)
out['device_per_component_s'] = min(pc)
out['device_speedup'] = out['device_per_component_s'] / (out['device_gemm_s'] + out['device_mon_s'])
print(json.dumps(out))
