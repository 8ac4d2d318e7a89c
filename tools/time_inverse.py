"""C5 workload: D-dimensional separable map (LET/iRBF/iRBF/RET + order-3 nonmonotone terms),
trained on N_train samples, then conditional sampling inverse_map(Z, X_star) with E conditioning columns.
Prints one JSON line with wall times through the public class API."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import synthetic_samples, c5_terms            # noqa: E402
from transport_map import transport_map                  # noqa: E402

D = int(os.environ.get('TTM_D', 256))
E = D // 2
ntrain = int(os.environ.get('TTM_NTRAIN', 10000))
ns = int(os.environ.get('TTM_NS', 1_250_000))
out = {'D': D, 'E': E, 'ntrain': ntrain, 'ns': ns}
X = synthetic_samples(ntrain, D, seed=0)
mon, non = c5_terms(D)
t = time.perf_counter()
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='separable monotonicity', verbose=False)
torch.cuda.synchronize()
out['ctor_s'] = time.perf_counter() - t
t = time.perf_counter()
tm.optimize()
torch.cuda.synchronize()
out['optimize_s'] = time.perf_counter() - t
rng = np.random.default_rng(1)
Xnew = synthetic_samples(ns, D, seed=2)
Z = rng.standard_normal((ns, D - E))
for mode, alt in (('table', True), ('bisect', False)):
    tm.alternate_root_finding = alt
    n_use = ns if alt else min(ns, int(os.environ.get('TTM_NS_BISECT', 200_000)))
    tm.inverse_map(Z[:1000], X_star=Xnew[:1000, :E])       # warm-up
    torch.cuda.synchronize()
    t = time.perf_counter()
    Xs = tm.inverse_map(Z[:n_use], X_star=Xnew[:n_use, :E])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    out['inverse_%s_s' % mode] = dt
    out['inverse_%s_samples_per_s' % mode] = n_use / dt
    if not alt:
        Zb = tm.map(Xs[:20000])[:, E:]
        out['bisect_roundtrip_max_abs'] = float(np.max(np.abs(Zb - Z[:20000])))
t = time.perf_counter()
Zm = tm.map(Xnew[:200_000])
torch.cuda.synchronize()
out['map_200k_s'] = time.perf_counter() - t
print(json.dumps(out))
