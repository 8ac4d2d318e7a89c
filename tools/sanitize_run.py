"""Small exercise of every round-2 kernel for compute-sanitizer (memcheck / racecheck):
tile K-objgrad (Gram and two-sweep forms, partial tiles), K-S, K-gram + tail, dense K-basis, K-inv-fused, K-inv-rect,
K-map-fused (map / pullback / pushforward), K-sepobj with the mapped result mirror."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from cases import synthetic_samples, c4_terms, c5_terms, ex05_terms, headline_sep_coeffs   # noqa: E402
from transport_map import transport_map                                                   # noqa: E402

rng = np.random.default_rng(0)
for gram in ('1', '0'):
    os.environ['TTM_GRAM'] = gram
    D, n = 12, 1000                                     # 3.9 tiles of 256: partial last tile, skipped warps
    mon, non = c4_terms(D)
    tm = transport_map(X=synthetic_samples(n, D, seed=1), monotone=mon, nonmonotone=non,
                       monotonicity='integrated rectifier', quadrature_input={'order': 10}, verbose=False)
    for k in (0, 1, 5, 11):
        c = rng.standard_normal(len(non[k]) + len(mon[k])) * 0.1
        f = tm.objective_function(c, k, len(non[k]))
        g = tm.objective_function_jacobian(c, k, len(non[k]))
        assert np.isfinite(f) and np.all(np.isfinite(g))
    tm.Psi_nonmon[11]
    Z = tm.map(synthetic_samples(300, D, seed=2))
    assert np.all(np.isfinite(Z))
del os.environ['TTM_GRAM']
D, E, n = 20, 3, 700
mon, non = c5_terms(D)
ts = transport_map(X=synthetic_samples(500, D, seed=3), monotone=mon, nonmonotone=non,
                   monotonicity='separable monotonicity', verbose=False)
ts.optimize(K=[0, 7, 19])
cm, cn = headline_sep_coeffs(mon, non)
for k in range(D):
    ts.coeffs_mon[k], ts.coeffs_nonmon[k] = cm[k].copy(), cn[k].copy()
X = ts.inverse_map(rng.standard_normal((n, D - E)), X_star=synthetic_samples(n, D, seed=4)[:, :E].copy())
assert np.all(np.isfinite(X))
# K-inv-rect (DMMA GEMM over the conditioning block) + the walk with staged tables: conditioning width that is no
# multiple of the 8-variable chunk, sample count that is no multiple of the 64-sample tile; 3- and 6-slot operands
os.environ['TTM_INV_SPLIT'] = '1'
for mixed in (False, True):
    D, E, n = 21, 9, 333
    mon, non = c5_terms(D)
    if mixed:
        non = [[[]] + [t for j in range(k) for t in ([j], [j, 'HF'], [j, j], [j, j, 'HF'], [j, j, j], [j, j, j, 'HF'])]
               for k in range(D)]
    tq = transport_map(X=synthetic_samples(400, D, seed=6), monotone=mon, nonmonotone=non,
                       monotonicity='separable monotonicity', verbose=False)
    cm, cn = headline_sep_coeffs(mon, non)
    for k in range(D):
        tq.coeffs_mon[k], tq.coeffs_nonmon[k] = cm[k].copy(), cn[k].copy()
    fz = tq._inverse_fused_setup([(i, k) for i, k in enumerate(range(E, D))])
    assert fz is not None and fz['R'] is not None
    X = tq.inverse_map(rng.standard_normal((n, D - E)), X_star=synthetic_samples(n, D, seed=7)[:, :E].copy())
    assert np.all(np.isfinite(X))
del os.environ['TTM_INV_SPLIT']
mon, non = ex05_terms()
t5 = transport_map(X=synthetic_samples(400, 2, seed=5), monotone=mon, nonmonotone=non,
                   monotonicity='separable monotonicity', verbose=False)
t5.optimize()
pts = rng.uniform(-2, 2, (333, 2))
d0 = t5.evaluate_pullback_density(pts)
d1 = t5.evaluate_pushforward_density(rng.standard_normal((333, 2)), lambda x: -0.5 * np.sum(x ** 2, axis=1))
assert np.all(np.isfinite(d0)) and np.all(np.isfinite(d1)) and np.all(np.isfinite(t5.map(pts)))
print('sanitize_run ok')
