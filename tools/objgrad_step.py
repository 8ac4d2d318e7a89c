"""One bench step (the 64 K-objgrad launches of the C4 map, N = 1M, Q = 100) after one warm-up step, for ncu:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:objgrad_tile -s 64 -c 64 \
        --csv --log-file gpurun_out/ncu_step.csv python tools/objgrad_step.py
    python tools/objgrad_step.py --summarise gpurun_out/ncu_step.csv profiles/ncu_traffic_r2.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def summarise(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if 'Metric Name' in r)
    ci = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= ci['Metric Value']:
            continue
        per.setdefault(r[ci['ID']], {})[r[ci['Metric Name']]] = (float(r[ci['Metric Value']].replace(',', '')), r[ci['Metric Unit']])
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    ids = sorted(per, key=int)
    rd = [per[i]['dram__bytes_read.sum'][0] * unit[per[i]['dram__bytes_read.sum'][1]] for i in ids]
    wr = [per[i]['dram__bytes_write.sum'][0] * unit[per[i]['dram__bytes_write.sum'][1]] for i in ids]
    fp = [per[i]['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][0] for i in ids]
    n = len(ids)
    alg = [8.0 * 1_000_000 * (k + 1) for k in range(n)]
    out = {'launches': n, 'per_launch_avg_bytes': (sum(rd) + sum(wr)) / n, 'read_avg_bytes': sum(rd) / n,
           'write_avg_bytes': sum(wr) / n, 'algorithmic_per_launch_avg_bytes': sum(alg) / n,
           'ratio_to_algorithmic': (sum(rd) + sum(wr)) / sum(alg),
           'k0_bytes': rd[0] + wr[0], 'k63_bytes': rd[-1] + wr[-1],
           'pipe_fp64_active_pct': {'avg': sum(fp) / n, 'k0': fp[0], 'k31': fp[n // 2 - 1], 'k63': fp[-1]},
           'source': '%s: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active over the %d '
                     'launches of one bench step (tools/objgrad_step.py, N=1M, Q=100, Gram mode)' % (os.path.basename(src), n)}
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--summarise':
        summarise(sys.argv[2], sys.argv[3])
        sys.exit(0)
    import numpy as np
    import torch
    from cases import synthetic_samples, c4_terms
    from transport_map import transport_map
    from ttt_b200 import binding as B
    D, n, Q = 64, 1_000_000, 100
    mon, non = c4_terms(D)
    tm = transport_map(X=synthetic_samples(n, D, seed=0), monotone=mon, nonmonotone=non,
                       monotonicity='integrated rectifier', quadrature_input={'order': Q}, verbose=False)
    rng = np.random.default_rng(0)
    coefs = [rng.standard_normal(len(non[k]) + len(mon[k])) * 0.05 for k in range(D)]
    for k in range(D):
        tm._gram_nonmon(k)
        tm._set_coeffs(k, coefs[k][:len(non[k])], coefs[k][len(non[k]):])
    Xp, ld = B.c_void_p(tm._Xt.data_ptr()), tm._Xt.shape[1]
    for rep in range(2):
        for k in range(D):
            B.check(tm._lib.ttm_objgrad_ir_launch(tm._plans[k], Xp, ld, n, tm._stream()))
        torch.cuda.synchronize()
