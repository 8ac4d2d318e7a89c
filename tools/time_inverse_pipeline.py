"""C5 conditional sampling (D=256, E=128, 1.25M samples) through the class API: wall time of repeated
inverse_map calls, one-shot against the chunked pipeline and a few chunk sizes.  Prints JSON lines."""
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

D, E, ns = 256, 128, int(os.environ.get('TTM_NS', 1_250_000))
mon, non = c5_terms(D)
tm = transport_map(X=synthetic_samples(10000, D, seed=0), monotone=mon, nonmonotone=non,
                   monotonicity='separable monotonicity', verbose=False)
tm.optimize()
rng = np.random.default_rng(1)
Xstar = synthetic_samples(ns, D, seed=2)[:, :E].copy()
Z = rng.standard_normal((ns, D - E))
ref = None
for label, pmin, chunk in (('one-shot', 10 ** 12, 0), ('pipeline', 1000, 163840), ('pipeline', 1000, 81920),
                           ('pipeline', 1000, 327680)):
    os.environ['TTM_INV_PIPELINE_MIN'] = str(pmin)
    if chunk:
        os.environ['TTM_INV_CHUNK'] = str(chunk)
    times = []
    for _ in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = tm.inverse_map(Z, X_star=Xstar)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t)
    if ref is None:
        ref = out.copy()
    print(json.dumps({'mode': label, 'chunk': chunk, 'seconds': [round(t, 4) for t in times],
                      'samples_per_s': round(ns / min(times)), 'identical_to_one_shot': bool(np.array_equal(out, ref))}),
          flush=True)
