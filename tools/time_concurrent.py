"""Experiment: throughput of the fused objective kernel when several components are evaluated CONCURRENTLY on
separate streams with a reduced grid per launch (TTM_OBJ_BPS blocks per SM), so that blocks of different
launches co-reside on an SM and their phases (issue-bound sweeps vs FP64-bound node loop) overlap."""
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

n, D, q = 1_000_000, 64, 100
nstreams = int(os.environ.get('TTM_STREAMS', 2))
X = synthetic_samples(n, D, seed=0)
mon, non = c4_terms(D)
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                   quadrature_input={'order': q}, verbose=False)
rng = np.random.default_rng(0)
coefs = [rng.standard_normal(len(non[k]) + len(mon[k])) * 0.05 for k in range(D)]
for k in range(D):
    tm._gram_nonmon(k)
    tm._set_coeffs(k, coefs[k][:len(non[k])], coefs[k][len(non[k]):])
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(nstreams)]
Xp, ld = B.c_void_p(tm._Xt.data_ptr()), tm._Xt.shape[1]
order = sorted(range(D), reverse=True)


def step():
    for i, k in enumerate(order):
        s = streams[i % nstreams]
        B.check(tm._lib.ttm_objgrad_ir_launch(tm._plans[k], Xp, ld, n, B.c_void_p(s.cuda_stream)))


for _ in range(2):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in streams:
        s.wait_event(e0)
    step()
    for s in streams:
        e = torch.cuda.Event()
        e.record(s)
        torch.cuda.current_stream().wait_event(e)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps({'streams': nstreams, 'bps': os.environ.get('TTM_OBJ_BPS', '4'), 'step_ms': float(np.median(ts)),
                  'evals_per_s': 64e3 / float(np.median(ts))}))
