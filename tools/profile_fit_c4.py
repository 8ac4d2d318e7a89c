"""cProfile of the host side of optimize() on the C4 map (single host thread)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import synthetic_samples, c4_terms            # noqa: E402
from transport_map import transport_map                  # noqa: E402

X = synthetic_samples(1_000_000, 64, seed=0)
mon, non = c4_terms(64)
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                   quadrature_input={'order': 100}, verbose=False, fit_threads=1)
pr = cProfile.Profile()
pr.enable()
t = time.perf_counter()
tm.optimize()
torch.cuda.synchronize()
print('optimize', time.perf_counter() - t)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18)
print(s.getvalue()[-3500:])
