"""cProfile of the C5 constructor and optimize() (D=256 separable map, 10^4 training samples), single fit thread."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                              # noqa: E402
from cases import synthetic_samples, c5_terms             # noqa: E402
from transport_map import transport_map                   # noqa: E402

D = 256
mon, non = c5_terms(D)
X = synthetic_samples(10000, D, seed=0)
transport_map(X=X[:, :4], monotone=mon[:4], nonmonotone=non[:4], monotonicity='separable monotonicity', verbose=False)


def report(pr, n=18):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(n)
    print(s.getvalue()[-3400:])


pr = cProfile.Profile()
pr.enable()
t = time.perf_counter()
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='separable monotonicity', verbose=False,
                   fit_threads=int(os.environ.get('TTM_FIT_THREADS', 1)))
print('ctor', time.perf_counter() - t)
pr.disable()
report(pr)
pr = cProfile.Profile()
pr.enable()
t = time.perf_counter()
tm.optimize()
torch.cuda.synchronize()
print('optimize', time.perf_counter() - t, getattr(tm, '_last_timing', None))
pr.disable()
report(pr, 26)
