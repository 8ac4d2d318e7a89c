import sys, cProfile, pstats, io, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from cases import synthetic_samples, c5_terms
from transport_map import transport_map
D = 256
mon, non = c5_terms(D)
tm = transport_map(X=synthetic_samples(10000, D, seed=0), monotone=mon, nonmonotone=non, monotonicity='separable monotonicity', verbose=False)
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter(); tm.optimize(); torch.cuda.synchronize(); print('optimize', time.perf_counter() - t)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[-3800:])
