"""End-to-end optimize() of the C4 map (D=64, N=1M, Q=100 by default) through the public class API.
Prints wall time, number of fused evaluations and the stationarity of the result (size-independent check)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from cases import synthetic_samples, c4_terms            # noqa: E402
from transport_map import transport_map                  # noqa: E402

n = int(os.environ.get('TTM_N', 1_000_000))
q = int(os.environ.get('TTM_Q', 100))
D = int(os.environ.get('TTM_D', 64))
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
X = synthetic_samples(n, D, seed=0)
mon, non = c4_terms(D)
t = time.perf_counter()
tm = transport_map(X=X, monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                   quadrature_input={'order': q}, verbose=False)
torch.cuda.synchronize()
ctor = time.perf_counter() - t
calls = {'n': 0}
orig = tm._objgrad


def counted(c, k):
    ent = tm._fg_cache.get(k)
    if ent is None or ent[0] != np.ascontiguousarray(c, dtype=np.float64).tobytes():
        calls['n'] += 1
    return orig(c, k)


tm._objgrad = counted
t = time.perf_counter()
tm.optimize()
torch.cuda.synchronize()
fit = time.perf_counter() - t
gmax, Js = 0.0, []
for k in range(D):
    c = np.concatenate((tm.coeffs_nonmon[k], tm.coeffs_mon[k]))
    div = len(tm.coeffs_nonmon[k])
    Js.append(tm.objective_function(c, k, div))
    gmax = max(gmax, float(np.max(np.abs(tm.objective_function_jacobian(c, k, div)))))
t = time.perf_counter()
Z = tm.map(X[:200000])
torch.cuda.synchronize()
tmap = time.perf_counter() - t
if rank == 0:
    print(json.dumps({'n': n, 'D': D, 'Q': q, 'gpus': world, 'ctor_s': ctor, 'optimize_s': fit,
                      'fused_evals_this_rank': calls['n'], 'timing_rank0': tm._last_timing, 'max_abs_grad_at_solution': gmax,
                      'sum_J': float(np.sum(Js)), 'map_200k_s': tmap,
                      'Z_mean_abs_max': float(np.max(np.abs(Z.mean(axis=0)))), 'Z_std_range': [float(Z.std(axis=0).min()), float(Z.std(axis=0).max())]}))
if world > 1:
    dist.destroy_process_group()
