"""C3: one Ensemble-Transport-Filter cycle of Example 06 -- reset -> optimize -> map ->
inverse_map(X_star) -- on 4-column ensembles, separable map with L2 regularisation (lambda = 0.05).
Times the CUDA class and (optionally) the CPU oracle on the same inputs and checks the posterior means."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch                                             # noqa: E402
from cases import ex06_terms, ex06_cycle_inputs          # noqa: E402
from transport_map import transport_map                  # noqa: E402
from ttm_oracle import OracleMap                         # noqa: E402

mon, non = ex06_terms(3)
out = []
for N in (500, 1000, 10000, 100000):
    dummy, cyc = ex06_cycle_inputs(N)
    kw = dict(monotone=mon, nonmonotone=non, monotonicity='separable monotonicity', regularization='l2',
              regularization_lambda=0.05, verbose=False)
    tm = transport_map(X=dummy.copy(), **kw)

    def cycle(m):
        m.reset(cyc.copy())
        m.optimize()
        Z = m.map(cyc.copy())
        return m.inverse_map(X_star=np.full((N, 1), 1.5), Z=Z)

    cycle(tm)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        post = cycle(tm)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    rec = {'N': N, 'gpu_cycle_ms': 1e3 * float(np.median(ts)), 'posterior_mean': post.mean(axis=0).tolist()}
    if N <= 10000 or os.environ.get('TTM_CPU_ALL'):
        om = OracleMap(X=dummy.copy(), **kw)
        cycle(om)
        t = time.perf_counter()
        po = cycle(om)
        rec['cpu_oracle_cycle_ms'] = 1e3 * (time.perf_counter() - t)
        rec['max_abs_diff_vs_oracle'] = float(np.max(np.abs(po - post)))
    out.append(rec)
print(json.dumps(out))
