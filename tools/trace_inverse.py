"""Timeline of one steady-state inverse_map call at C5 (torch.profiler / CUPTI: memcpys and kernels with timestamps)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch                                             # noqa: E402
from torch.profiler import profile, ProfilerActivity     # noqa: E402
from cases import synthetic_samples, c5_terms, headline_sep_coeffs   # noqa: E402
from transport_map import transport_map                  # noqa: E402

D, E, ns = 256, 128, int(os.environ.get('TTM_NS', 1_250_000))
mon, non = c5_terms(D)
tm = transport_map(X=synthetic_samples(4000, D, seed=0), monotone=mon, nonmonotone=non,
                   monotonicity='separable monotonicity', verbose=False)
cm, cn = headline_sep_coeffs(mon, non)
for k in range(D):
    tm.coeffs_mon[k], tm.coeffs_nonmon[k] = cm[k].copy(), cn[k].copy()
rng = np.random.default_rng(1)
Xstar = torch.empty((ns, E), dtype=torch.float64, pin_memory=True).numpy()
Z = torch.empty((ns, D - E), dtype=torch.float64, pin_memory=True).numpy()
Xstar[:] = synthetic_samples(ns, D, seed=2)[:, :E]
Z[:] = rng.standard_normal((ns, D - E))
for _ in range(2):
    out = tm.inverse_map(Z, X_star=Xstar)
    del out
torch.cuda.synchronize()
t = time.perf_counter()
out = tm.inverse_map(Z, X_star=Xstar)
torch.cuda.synchronize()
print('untraced call s', time.perf_counter() - t)
del out
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t = time.perf_counter()
    out = tm.inverse_map(Z, X_star=Xstar)
    torch.cuda.synchronize()
    print('traced call s', time.perf_counter() - t)
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append({'name': e.name[:60], 'start_us': e.time_range.start, 'dur_us': e.time_range.end - e.time_range.start})
t0 = min(e['start_us'] for e in ev)
for e in ev:
    e['start_us'] -= t0
ev.sort(key=lambda e: e['start_us'])
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(ev, open(os.path.join(ROOT, 'gpurun_out', 'trace_inverse.json'), 'w'))
cpu = [(e.name[:50], e.time_range.start - t0, e.time_range.end - e.time_range.start) for e in prof.events()
       if e.device_type == torch.autograd.DeviceType.CPU and (e.time_range.end - e.time_range.start) > 500]
cpu.sort(key=lambda r: r[1])
json.dump(cpu, open(os.path.join(ROOT, 'gpurun_out', 'trace_inverse_cpu.json'), 'w'))
print('device events', len(ev), 'span ms', (max(e['start_us'] + e['dur_us'] for e in ev)) / 1e3)

# host-side view of the same call
import cProfile
import io
import pstats
pr = cProfile.Profile()
pr.enable()
out = tm.inverse_map(Z, X_star=Xstar)
torch.cuda.synchronize()
pr.disable()
sio = io.StringIO()
pstats.Stats(pr, stream=sio).sort_stats('cumulative').print_stats(25)
print(sio.getvalue()[-4500:])
# new coefficients every call (an EnTF-like loop): the operands are rebuilt
ts = []
for rep in range(3):
    for k in range(D):
        tm.coeffs_nonmon[k] = tm.coeffs_nonmon[k] * 1.0001
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = tm.inverse_map(Z, X_star=Xstar)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t)
print('call with new coefficients s', ts)
