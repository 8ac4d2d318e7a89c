"""Device-time the fused objective+gradient kernel for a few components of the C4 workload; prints one JSON line.
TTM_KERNEL=general times the general kernel instead of the tile kernel, TTM_GRAM=0 the two-sweep form."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import synthetic_samples, c4_terms            # noqa: E402
from transport_map import transport_map                  # noqa: E402
from ttt_b200 import binding as B                        # noqa: E402

n = int(os.environ.get('TTM_N', 1_000_000))
q = int(os.environ.get('TTM_Q', 100))
D = 64
X = synthetic_samples(n, D, seed=0)
mon, non = c4_terms(D)
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                   quadrature_input={'order': q}, verbose=False)
rng = np.random.default_rng(0)
coefs = [rng.standard_normal(len(non[k]) + len(mon[k])) * 0.05 for k in range(D)]
general = os.environ.get('TTM_KERNEL', 'tile') == 'general'
B.check(tm._lib.ttm_ctx_set_objgrad_kernel(tm._ctx, 1 if general else 0))
bps = int(os.environ.get('TTM_BPS', 0))
B.check(tm._lib.ttm_ctx_set_blocks_per_sm(tm._ctx, bps))
out = {'kernel': 'general' if general else 'tile', 'gram': os.environ.get('TTM_GRAM', 'auto'), 'bps': bps, 'n': n, 'q': q}
stream = tm._stream()
Xp, ld = B.c_void_p(tm._Xt.data_ptr()), tm._Xt.shape[1]
for k in (0, 1, 31, 63):
    out['k%d_gram' % k] = tm._gram_nonmon(k) is not None
    tm._set_coeffs(k, coefs[k][:len(non[k])], coefs[k][len(non[k]):])
    for _ in range(3):
        B.check(tm._lib.ttm_objgrad_ir_launch(tm._plans[k], Xp, ld, n, stream))
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.check(tm._lib.ttm_objgrad_ir_launch(tm._plans[k], Xp, ld, n, stream))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = np.empty(1 + len(coefs[k]))
    B.check(tm._lib.ttm_plan_get_out(tm._plans[k], B.dptr(res), res.size, stream))
    out['k%d_ms' % k] = round(float(np.median(ts)), 4)
    out['k%d_J' % k] = repr(float(res[0]))
    out['k%d_gsum' % k] = repr(float(np.abs(res[1:]).sum()))
print(json.dumps(out))
