"""Device-time the secondary kernels (K-std, K-basis, K-gram, K-sepobj, K-sep-eval, K-inv-table, K-inv-bisect)
on the C5-pattern separable map and report achieved GB/s / TFLOP/s against the algorithmic work of
SURVEY.md 8(d).  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import synthetic_samples, c5_terms            # noqa: E402
from transport_map import transport_map                  # noqa: E402
from ttt_b200 import binding as B                        # noqa: E402

D = int(os.environ.get('TTM_D', 64))
N = int(os.environ.get('TTM_N', 1_000_000))
X = synthetic_samples(N, D, seed=0)
mon, non = c5_terms(D)
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='separable monotonicity', verbose=False)
lib, st = tm._lib, tm._stream()
Xp, ld = B.c_void_p(tm._Xt.data_ptr()), tm._Xt.shape[1]


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


out = {'D': D, 'N': N}
k = D - 1
p = tm._host_plans[k]
M = p.m_non + p.m_mon
rng = np.random.default_rng(0)
tm._set_coeffs(k, rng.standard_normal(p.m_non) * 0.1, np.abs(rng.standard_normal(p.m_mon)) + 0.1)

# K-std (colstats + standardise/transpose) on the raw matrix
Xd = tm._upload(X)
mean, std, Xt2 = tm._empty(D), tm._empty(D), tm._empty(D, N)
t = timeit(lambda: (B.check(lib.ttm_colstats(tm._ctx, B.c_void_p(Xd.data_ptr()), N, D, B.c_void_p(mean.data_ptr()),
                                              B.c_void_p(std.data_ptr()), B.c_void_p(tm._scratch.data_ptr()), st)),
                    B.check(lib.ttm_standardize_transpose(tm._ctx, B.c_void_p(Xd.data_ptr()), N, D,
                                                          B.c_void_p(mean.data_ptr()), B.c_void_p(std.data_ptr()),
                                                          B.c_void_p(Xt2.data_ptr()), N, st))))
out['std'] = {'ms': t * 1e3, 'GBps': 4 * 8 * N * D / t / 1e9, 'bytes': '3 reads + 1 write of 8*N*D'}
del Xd, Xt2

# K-basis (nonmonotone matrix of the last component) on a slice
nb = min(N, 200_000)
Psi = tm._empty(nb, p.m_non)
t = timeit(lambda: B.check(lib.ttm_basis_eval(tm._plans[k], 0, Xp, ld, nb, B.c_void_p(Psi.data_ptr()), st)))
out['basis'] = {'ms': t * 1e3, 'GBps': 8 * nb * (k + p.m_non) / t / 1e9, 'n': nb, 'm': p.m_non}
del Psi

# K-sep-eval (S and dS of the last component)
S, dS = tm._empty(N), tm._empty(N)
t = timeit(lambda: B.check(lib.ttm_sep_eval(tm._plans[k], Xp, ld, N, B.c_void_p(S.data_ptr()), Xp, ld,
                                            B.c_void_p(dS.data_ptr()), st)))
out['sep_eval'] = {'ms': t * 1e3, 'GBps': 8 * N * (k + 1 + 2) / t / 1e9}

# K-sepobj
b = np.abs(rng.standard_normal(p.m_mon)) + 0.1
res = np.empty(1 + p.m_dmon)
t = timeit(lambda: B.check(lib.ttm_sep_objgrad(tm._plans[k], Xp, ld, N, B.dptr(b), B.dptr(res), st)))
out['sepobj'] = {'ms': t * 1e3, 'GBps': 8 * N / t / 1e9, 'GFLOPs': 190.0 * N / t / 1e9, 'note': 'includes the D2H of (1+m_mon) doubles'}

# K-gram on a training-sized slice
ng = int(os.environ.get('TTM_NGRAM', 100_000))
G = tm._empty(M, M)
Mp = (M + 7) // 8 * 8
need = 64 * Mp * Mp + min(ng, 1 << 18) * Mp
if tm._scratch.numel() < need:
    tm._scratch = tm._empty(need)
t = timeit(lambda: B.check(lib.ttm_gram(tm._plans[k], Xp, ld, ng, B.c_void_p(G.data_ptr()),
                                        B.c_void_p(tm._scratch.data_ptr()), tm._scratch.numel(), st)), reps=3)
out['gram'] = {'ms': t * 1e3, 'TFLOPs': 2.0 * ng * M * M / t / 1e12, 'n': ng, 'M': M}
out['gram_sizes'] = []
for ng2 in (10_000, 200_000):
    if ng2 <= N:
        t2 = timeit(lambda: B.check(lib.ttm_gram(tm._plans[k], Xp, ld, ng2, B.c_void_p(G.data_ptr()),
                                                 B.c_void_p(tm._scratch.data_ptr()), tm._scratch.numel(), st)), reps=3)
        out['gram_sizes'].append({'n': ng2, 'M': M, 'ms': t2 * 1e3, 'TFLOPs': 2.0 * ng2 * M * M / t2 / 1e12})
tt = timeit(lambda: B.check(lib.ttm_gram_tail(tm._plans[k], Xp, ld, ng, p.m_non, B.c_void_p(G.data_ptr()),
                                              B.c_void_p(tm._scratch.data_ptr()), tm._scratch.numel(), st)), reps=3)
out['gram_tail'] = {'ms': tt * 1e3, 'n': ng, 'M': M, 'first_col': p.m_non}

# K-inv-table / K-inv-bisect for the last component (columns < k solved = the training columns)
z = tm._upload(rng.standard_normal(N))
Xw = tm._Xt.clone()
pts = np.linspace(-10, 10, 1001)
tab = tm._empty(2002)
tab[1001:] = tm._upload(pts)
B.check(lib.ttm_mon_table(tm._plans[k], 1001, B.c_void_p(tab.data_ptr()), st))
o = tab[:1001].cpu().numpy()
ind = np.argsort(o, kind='mergesort')
tab = tm._upload(np.concatenate((o[ind], pts[ind])))
t = timeit(lambda: B.check(lib.ttm_inverse_table(tm._plans[k], B.c_void_p(Xw.data_ptr()), Xw.shape[1], N,
                                                 B.c_void_p(z.data_ptr()), B.c_void_p(tab.data_ptr()), 1001, 1, st)))
out['inverse_table'] = {'ms': t * 1e3, 'GBps': 8 * N * (k + 2) / t / 1e9, 'Msamples_per_s': N / t / 1e6}
stalled = B.c_int(0)
t = timeit(lambda: B.check(lib.ttm_inverse_bisect(tm._plans[k], B.c_void_p(Xw.data_ptr()), Xw.shape[1], N,
                                                  B.c_void_p(z.data_ptr()), 1, 100, B.ctypes.byref(stalled), st)), reps=3)
out['inverse_bisect'] = {'ms': t * 1e3, 'Msamples_per_s': N / t / 1e6}
print(json.dumps(out))
