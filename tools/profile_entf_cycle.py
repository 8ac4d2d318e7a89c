"""cProfile of one Example-06 EnTF cycle (reset -> optimize -> map -> inverse_map) at N=1000."""
import cProfile
import io
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import ex06_terms, ex06_cycle_inputs          # noqa: E402
from transport_map import transport_map                  # noqa: E402

N = int(os.environ.get('TTM_N', 1000))
mon, non = ex06_terms(3)
dummy, cyc = ex06_cycle_inputs(N)
tm = transport_map(X=dummy.copy(), monotone=mon, nonmonotone=non, monotonicity='separable monotonicity',
                   regularization='l2', regularization_lambda=0.05, verbose=False)


def cycle():
    tm.reset(cyc.copy())
    tm.optimize()
    Z = tm.map(cyc.copy())
    return tm.inverse_map(X_star=np.full((N, 1), 1.5), Z=Z)


for _ in range(3):
    cycle()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    cycle()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[-7000:])
