"""
Generate the golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, numpy 2.3.5, scipy 1.18.1):

    python tests/golden/make_golden.py

Writes tests/golden/<case>.npz (outputs of tests/harness.run_case on
`/root/reference/transport_map.py`) and tests/golden/ex01_known_answer.npz
(the Example-01/02 shipped coefficient pickles with the regenerated spiral
ensemble, objective values and gradients).  The GPU box never runs this script;
it only reads the committed .npz files.
"""

import copy
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, '/root/reference')

import scipy                                                     # noqa: E402
from transport_map import transport_map                          # noqa: E402  (the reference)
from cases import cases, ex01_terms                              # noqa: E402
from harness import run_case                                     # noqa: E402

EX01 = '/root/reference/Examples A - spiral distribution/Example 01 - full map/'
EX02 = '/root/reference/Examples A - spiral distribution/Example 02 - partial map/'


def make_reference(X, **kw):
    return transport_map(X=X, **kw)


def spiral(size):
    """Input recipe of example_01.py:31-57 (formula restated; draws use the global numpy seed)."""
    import scipy.stats
    seeds = scipy.stats.beta.rvs(a=4, b=3, size=size) * 3 * np.pi - np.pi
    vals = (seeds + np.pi) / (3 * np.pi) * 6 - 3
    X = np.column_stack((np.cos(seeds), np.sin(seeds))) * ((1 + seeds + np.pi) / (3 * np.pi) * 5)[:, None]
    X += np.column_stack((np.cos(seeds), np.sin(seeds))) * \
        (scipy.stats.norm.rvs(size=size) * scipy.stats.norm.pdf(vals))[:, None]
    return X / 2


def known_answer():
    np.random.seed(0)
    X = spiral(10000)
    mon, non = ex01_terms(10)
    qi = {'order': 25, 'adaptive': False, 'threshold': 1e-9, 'verbose': False, 'increment': 6}
    out = {'X': X}
    d = pickle.load(open(EX01 + 'dict_coeffs_order=10.p', 'rb'))
    tm = transport_map(X=copy.copy(X), monotone=mon, nonmonotone=non, verbose=False,
                       monotonicity='integrated rectifier', quadrature_input=dict(qi))
    tm.coeffs_mon, tm.coeffs_nonmon = copy.copy(d['coeffs_mon']), copy.copy(d['coeffs_nonmon'])
    for k in range(2):
        c = np.concatenate((tm.coeffs_nonmon[k], tm.coeffs_mon[k]))
        div = len(tm.coeffs_nonmon[k])
        out['full_coeffs_%d' % k], out['full_div_%d' % k] = c, np.asarray(div)
        out['full_J_%d' % k] = np.asarray(tm.objective_function(c.copy(), k, div))
        out['full_grad_%d' % k] = tm.objective_function_jacobian(c.copy(), k, div)
    out['full_map_head'] = tm.map(copy.copy(X))[:512]
    out['full_map_std'] = tm.map(copy.copy(X)).std(axis=0)
    d2 = pickle.load(open(EX02 + 'dict_coeffs_order=10_partial.p', 'rb'))
    tm2 = transport_map(X=copy.copy(X), monotone=mon[1:], nonmonotone=non[1:], verbose=False,
                        monotonicity='integrated rectifier', quadrature_input=dict(qi))
    tm2.coeffs_mon, tm2.coeffs_nonmon = copy.copy(d2['coeffs_mon']), copy.copy(d2['coeffs_nonmon'])
    c = np.concatenate((tm2.coeffs_nonmon[0], tm2.coeffs_mon[0]))
    div = len(tm2.coeffs_nonmon[0])
    out['partial_coeffs_0'], out['partial_div_0'] = c, np.asarray(div)
    out['partial_J_0'] = np.asarray(tm2.objective_function(c.copy(), 0, div))
    out['partial_grad_0'] = tm2.objective_function_jacobian(c.copy(), 0, div)
    print('known answers: J0=%r J1=%r |g0|=%.2e |g1|=%.2e  partial J=%r |g|=%.2e' % (
        float(out['full_J_0']), float(out['full_J_1']), np.linalg.norm(out['full_grad_0']),
        np.linalg.norm(out['full_grad_1']), float(out['partial_J_0']), np.linalg.norm(out['partial_grad_0'])))
    np.savez_compressed(os.path.join(HERE, 'ex01_known_answer.npz'), **out)


def main():
    print('numpy', np.__version__, 'scipy', scipy.__version__)
    for name, case in cases().items():
        res = run_case(make_reference, case)
        res['_versions'] = np.asarray('numpy %s scipy %s' % (np.__version__, scipy.__version__))
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)
        print('%-40s %3d arrays' % (name, len(res)))
    known_answer()


if __name__ == '__main__':
    main()
