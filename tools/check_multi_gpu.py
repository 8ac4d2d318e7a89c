"""Run under torchrun with >= 2 GPUs: checks both multi-GPU modes of optimize() against the single-GPU fit.
  (a) component-sharded: ranks fit disjoint components, one all-gather -> identical coefficients everywhere;
  (b) sample-sharded:    ranks hold disjoint sample shards, all-reduce of (J, grad) / Gram per evaluation."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
import torch.distributed as dist                         # noqa: E402
from cases import synthetic_samples, c4_terms, ex06_terms  # noqa: E402
from transport_map import transport_map                  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1, np.abs(b))))


N, D = 40000, 6
X = synthetic_samples(N, D, seed=5)
mon, non = c4_terms(D)
kw = dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier', verbose=False)
# (a) component sharded (default when torch.distributed is initialised)
tm = transport_map(X=X.copy(), quadrature_input={'order': 25}, **kw)
tm.optimize()
flat = np.concatenate([np.concatenate((tm.coeffs_nonmon[k], tm.coeffs_mon[k])) for k in range(D)])
t = torch.from_numpy(flat).cuda()
ref = t.clone()
dist.broadcast(ref, 0)
assert torch.equal(t, ref), 'coefficients differ between ranks after the all-gather'
# single-GPU reference fit of the same map on this rank only (no sharding): force world of one
single = transport_map(X=X.copy(), quadrature_input={'order': 25}, **kw)
single._world, single._sharded = 1, False
import ttt_b200.parallel as par                          # noqa: E402
_world = par.world
par.world = lambda: (0, 1)
single.optimize()
par.world = _world
flat1 = np.concatenate([np.concatenate((single.coeffs_nonmon[k], single.coeffs_mon[k])) for k in range(D)])
e_comp = rel(flat, flat1)
# (b) sample sharded
lo, hi = rank * N // world, (rank + 1) * N // world
ts = transport_map(X=X[lo:hi].copy(), quadrature_input={'order': 25}, sample_sharded=True, **kw)
assert rel(ts.X_mean, single.X_mean) < 1e-12 and rel(ts.X_std, single.X_std) < 1e-12
c = np.random.default_rng(3).standard_normal(len(non[3]) + len(mon[3])) * 0.1
e_obj = abs(ts.objective_function(c, 3, len(non[3])) - single.objective_function(c, 3, len(non[3])))
e_grad = rel(ts.objective_function_jacobian(c, 3, len(non[3])), single.objective_function_jacobian(c, 3, len(non[3])))
ts.optimize()
flat2 = np.concatenate([np.concatenate((ts.coeffs_nonmon[k], ts.coeffs_mon[k])) for k in range(D)])
e_samp = rel(flat2, flat1)
# (c) map() / inverse_map() shard by samples with no exchange: every rank maps its own rows
for k in range(D):
    ts.coeffs_nonmon[k], ts.coeffs_mon[k] = single.coeffs_nonmon[k].copy(), single.coeffs_mon[k].copy()
Zfull = single.map(X.copy())
e_map = rel(ts.map(X[lo:hi].copy()), Zfull[lo:hi])
Zr = np.random.default_rng(11).standard_normal((2048, D))
r0, r1 = rank * 2048 // world, (rank + 1) * 2048 // world
e_inv = float(np.max(np.abs(ts.inverse_map(Zr[r0:r1].copy()) - single.inverse_map(Zr.copy())[r0:r1])))
# separable + L2 (Gram all-reduce)
mon6, non6 = ex06_terms(3)
X6 = np.column_stack((X[:, 0] + 0.3 * X[:, 1], X[:, :3]))
kw6 = dict(monotone=mon6, nonmonotone=non6, monotonicity='separable monotonicity', regularization='l2',
           regularization_lambda=0.05, verbose=False)
s1 = transport_map(X=X6.copy(), **kw6)
par.world = lambda: (0, 1)
s1._world, s1._sharded = 1, False
s1.optimize()
par.world = _world
s2 = transport_map(X=X6[lo:hi].copy(), sample_sharded=True, **kw6)
s2.optimize()
e_sep = max(rel(s2.coeffs_mon[k], s1.coeffs_mon[k]) for k in range(3))
e_sep = max(e_sep, max(rel(s2.coeffs_nonmon[k], s1.coeffs_nonmon[k]) for k in range(3)))
# pullback density, sample-sharded (fused small-map kernel) against the whole-ensemble map
for k in range(3):
    s2.coeffs_nonmon[k], s2.coeffs_mon[k] = s1.coeffs_nonmon[k].copy(), s1.coeffs_mon[k].copy()
e_dens = rel(s2.evaluate_pullback_density(X6[lo:hi, 1:].copy(), X_star=X6[lo:hi, :1].copy()),
             s1.evaluate_pullback_density(X6[:, 1:].copy(), X_star=X6[:, :1].copy())[lo:hi])
errs = torch.tensor([e_comp, e_obj, e_grad, e_samp, e_sep, e_map, e_inv, e_dens], dtype=torch.float64, device='cuda')
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
e_comp, e_obj, e_grad, e_samp, e_sep, e_map, e_inv, e_dens = [float(v) for v in errs]
if rank == 0:
    line = ('MULTI_GPU_CHECK world=%d component_sharded=%.2e sample_sharded: obj=%.2e grad=%.2e coeffs=%.2e separable=%.2e '
            'map=%.2e inverse=%.2e pullback=%.2e' % (world, e_comp, e_obj, e_grad, e_samp, e_sep, e_map, e_inv, e_dens))
    print(line)
    out = os.environ.get('TTM_MULTI_GPU_LOG')
    if out:
        with open(out, 'a') as f:
            f.write(line + '\n')
assert e_comp == 0.0 and e_obj < 1e-12 and e_grad < 1e-11 and e_samp < 1e-6 and e_sep < 1e-6
assert e_map < 1e-9 and e_inv < 2e-8 and e_dens < 1e-9
dist.destroy_process_group()
