"""
Shared parity workloads (map specifications + seeded inputs).

Used by tests/golden/make_golden.py (run against the unmodified reference in the
build container), by the CPU oracle tests and by the GPU parity tests, so that
all three see exactly the same constructor arguments.
Term-list recipes follow SURVEY.md section 8(d) and the reference examples
(example_01.py:121-170, example_05.py:80-112, example_06.py:202-214).
"""

import copy
import itertools

import numpy as np


def synthetic_samples(N, D, seed=0):
    """Weakly non-Gaussian Markov chain of SURVEY.md 8(d)."""
    rng = np.random.default_rng(seed)
    Zs = rng.standard_normal((N, D))
    X = np.zeros((N, D))
    X[:, 0] = Zs[:, 0]
    for d in range(1, D):
        X[:, d] = 0.6 * X[:, d - 1] + 0.3 * np.tanh(X[:, d - 1]) ** 2 + 0.8 * Zs[:, d]
    return X


def c4_terms(D):
    """C4: order-3 Hermite-function integrated-rectifier map."""
    mon, non = [], []
    for k in range(D):
        mon.append([[k, 'HF'], [k, k, 'HF'], [k, k, k, 'HF']] + ([[k - 1, k, 'HF']] if k > 0 else []))
        n = [[]]
        for j in range(k):
            n += [[j], [j, j, 'HF'], [j, j, j, 'HF']]
        non.append(n)
    return mon, non


def c5_terms(D):
    """C5: separable map with edge terms + iRBFs (example_06.py:211-214 pattern)."""
    mon, non = [], []
    for k in range(D):
        mon.append(['LET %d' % k, 'iRBF %d' % k, 'iRBF %d' % k, 'RET %d' % k])
        n = [[]]
        for j in range(k):
            n += [[j], [j, j, 'HF'], [j, j, j, 'HF']]
        non.append(n)
    return mon, non


def ex01_terms(maxorder):
    """Example 01 (spiral) term lists, example_01.py:121-170."""
    mon, non = [], []
    for k in range(2):
        mon.append([])
        non.append([[]])
        for order in range(maxorder):
            if k > 0:
                non[-1].append([k - 1] * (order + 1) + ['HF'])
            for entry in itertools.combinations_with_replacement(range(k + 1), order + 1):
                if k in entry:
                    mon[-1].append([int(e) for e in entry] + ['HF'])
    return mon, non


def ex05_terms():
    """Example 05 (densities) term lists, example_05.py:80-112."""
    mon = [[[0], 'iRBF 0', 'iRBF 0'], [[1], 'iRBF 1', 'iRBF 1']]
    non = [[[]], [[], [0], [0, 0, 'HF'], [0, 0, 0, 'HF']]]
    return mon, non


def ex06_terms(order=3):
    """Example 06 (EnTF) nonlinear filter map, example_06.py:202-214."""
    non = [
        [[], [0]] + [[0] * od + ['HF'] for od in range(1, order + 1)],
        [[], [1]] + [[1] * od + ['HF'] for od in range(1, order + 1)],
        [[], [1]] + [[1] * od + ['HF'] for od in range(1, order + 1)] + [[2]]
        + [[2] * od + ['HF'] for od in range(1, order + 1)]]
    mon = [['LET 1'] + ['iRBF 1'] * (order - 1) + ['RET 1'], [[2]], [[3]]]
    return mon, non


def ex06_cycle_inputs(N):
    """One EnTF cycle input of SURVEY.md 8(d) C3 (needs scipy.stats + global numpy seed)."""
    import scipy.stats
    np.random.seed(0)
    dummy = np.random.uniform(size=(N, 4))
    Xs = scipy.stats.norm.rvs(size=(N, 3)) * np.array([8, 9, 8]) + np.array([0, 0, 25])
    Xs[:, 1] += 0.8 * Xs[:, 0]
    Yt = Xs[:, 0] + scipy.stats.norm.rvs(scale=2, size=N)
    return dummy, np.column_stack((Yt[:, None], Xs))


FAMILIES = ['power series', 'hermite', 'hermite_e', 'chebyshev', 'laguerre', 'legendre', 'hermite function']


def family_terms():
    """A deliberately mixed bag: plain/HF factors, products, all four special terms."""
    mon = [[[0], [0, 0, 'HF'], 'RBF 0'],
           [[1], [0, 1], [0, 0, 1, 1, 'HF'], 'LET 1', 'RET 1'],
           [[2], [2, 2, 2], [2, 2, 2, 'HF'], [1, 2, 2], 'iRBF 2', 'RBF 2', 'iRBF 2']]
    non = [[[]],
           [[], [0], [0, 0, 0, 0], 'RBF 0', 'RBF 0', 'RBF 0'],
           [[], [0, 1], [0, 0, 1, 'HF'], 'LET 0', 'iRBF 1', 'RET 0']]
    return mon, non


def cases():
    """name -> dict(kwargs=<ctor kwargs without X>, X=<training samples>, extra...)."""
    out = {}

    mon, non = c4_terms(4)
    out['ir_c4_d4'] = dict(
        X=synthetic_samples(256, 4, seed=0),
        kwargs=dict(monotone=mon, nonmonotone=non, polynomial_type='hermite function',
                    monotonicity='integrated rectifier', quadrature_input={'order': 25}),
        fit=True, n_inverse=48)

    mon, non = c4_terms(3)
    out['ir_c4_q100_delta0'] = dict(
        X=synthetic_samples(128, 3, seed=1) * np.array([2.0, 0.5, 3.0]) + np.array([1.0, -2.0, 0.3]),
        kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier', delta=0.0),
        fit=False, n_inverse=0)

    mon, non = c4_terms(3)
    lam = [np.linspace(0.01, 0.2, len(non[k]) + len(mon[k])) for k in range(3)]
    for rect in ['softplus', 'expneg']:
        for reg, l in [('l1', 0.05), ('l2', lam)]:
            out['ir_%s_%s' % (rect, reg)] = dict(
                X=synthetic_samples(128, 3, seed=2),
                kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                            quadrature_input={'order': 20}, rectifier_type=rect,
                            regularization=reg, regularization_lambda=l, delta=1e-6),
                fit=False, n_inverse=16 if rect == 'softplus' else 0)

    mon, non = ex01_terms(4)
    out['ir_ex01_order4'] = dict(
        X=synthetic_samples(200, 2, seed=3) * np.array([1.0, 1.7]),
        kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                    quadrature_input={'order': 25}),
        fit=False, n_inverse=16)

    mon, non = ex01_terms(4)
    out['ir_ex01_partial'] = dict(
        X=synthetic_samples(200, 2, seed=4),
        kwargs=dict(monotone=mon[1:], nonmonotone=non[1:], monotonicity='integrated rectifier',
                    quadrature_input={'order': 15}),
        fit=True, n_inverse=16)

    mon, non = ex05_terms()
    X5 = synthetic_samples(1000, 2, seed=5) * np.array([2.0, 0.5]) + np.array([1.0, -3.0])
    out['sep_ex05'] = dict(
        X=X5, kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='separable monotonicity'),
        fit=True, n_inverse=200, densities=True)
    out['sep_ex05_partial_l2'] = dict(
        X=X5, kwargs=dict(monotone=mon[1:], nonmonotone=non[1:], monotonicity='separable monotonicity',
                          regularization='l2', regularization_lambda=0.05),
        fit=True, n_inverse=200, densities=True)

    mon, non = ex06_terms(3)
    dummy, cyc = ex06_cycle_inputs(500)
    out['sep_ex06_cycle'] = dict(
        X=dummy, reset_X=cyc,
        kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='separable monotonicity',
                    regularization='l2', regularization_lambda=0.05),
        fit=True, n_inverse=500, ystar=1.5)

    mon, non = c5_terms(6)
    out['sep_c5_d6'] = dict(
        X=synthetic_samples(400, 6, seed=6),
        kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='separable monotonicity'),
        fit=True, n_inverse=100, cond=3)

    mon, non = family_terms()
    for fam in FAMILIES:
        for mono in ['integrated rectifier', 'separable monotonicity']:
            tag = 'fam_%s_%s' % (fam.replace(' ', '_'), 'ir' if mono.startswith('int') else 'sep')
            out[tag] = dict(
                X=synthetic_samples(64, 3, seed=7),
                kwargs=dict(monotone=mon, nonmonotone=non, polynomial_type=fam, monotonicity=mono,
                            quadrature_input={'order': 10}, ST_scale_factor=1.3),
                fit=False, n_inverse=0, objgrad=mono.startswith('int'))

    mon = [[[0]], [[1], 'RBF 0', 'RBF 0', 'iRBF 1', 'iRBF 1', 'iRBF 1'], [[2], 'RBF 0', 'RBF 1', 'iRBF 2', 'iRBF 2']]
    non = [[[]], [[], [0]], [[], [0], [1]]]
    out['ir_cross_st_quantile'] = dict(
        X=synthetic_samples(64, 3, seed=8),
        kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                    quadrature_input={'order': 10}, standardization='quantiles', ST_scale_mode='static'),
        fit=False, n_inverse=8)

    mon = [[[0], [0, 'HF'], [0, 0, 'HF'], [0, 0]], [[1, 'HF'], [1], [0, 1, 'HF'], [0, 1]]]
    non = [[[]], [[], [0]]]
    out['sep_der_key_collision'] = dict(
        X=synthetic_samples(64, 2, seed=9),
        kwargs=dict(monotone=mon, nonmonotone=non, monotonicity='separable monotonicity',
                    standardize_samples=False),
        fit=False, n_inverse=0)

    for c in out.values():
        c['kwargs'] = copy.deepcopy(c['kwargs'])
        c['kwargs']['verbose'] = False
    return out


# ---- headline shapes (BASELINE.json configs C4 / C5 at their real D), fixtures from tests/golden/make_golden_headline.py
HEADLINE_N = 4000
HEADLINE_KS = (0, 1, 5, 31, 63)
HEADLINE_QS = (25, 100)
C5_INV = dict(D=256, E=128, ntrain=1000, n_table=500, n_bisect=200, seed=4242)


def headline_coeffs(mon, non, k):
    """Seeded coefficient vector [nonmonotone | monotone] of component k for the C4 objective/gradient checks."""
    rng = np.random.default_rng(7000 + k)
    return rng.standard_normal(len(non[k]) + len(mon[k])) * 0.05


def headline_sep_coeffs(mon, non):
    """Seeded coefficients of a separable map: nonmonotone N(0, 0.1^2)/sqrt(1+k), monotone positive (0.05 .. 0.55)."""
    rng = np.random.default_rng(7100)
    cm = [np.abs(rng.standard_normal(len(mon[k]))) * 0.25 + 0.05 for k in range(len(mon))]
    cn = [rng.standard_normal(len(non[k])) * 0.1 / np.sqrt(1.0 + k) for k in range(len(mon))]
    return cm, cn


def adapt_separable_case():
    """Seeded input of the separable structure search (adapt_map, tm.py:373-657): a 3-D ensemble with one skewed
    marginal and one nonlinear dependence."""
    rng = np.random.default_rng(31)
    n = 400
    x0 = rng.standard_normal(n)
    x1 = np.exp(0.5 * rng.standard_normal(n)) + 0.2 * x0
    x2 = 0.7 * x0 ** 2 + 0.5 * rng.standard_normal(n)
    X = np.column_stack((x0, x1, x2))
    kw = dict(monotone=None, nonmonotone=None, monotonicity='separable monotonicity', adaptation=True,
              adaptation_map_type='separable', verbose=False)
    call = dict(maxorder_mon=4, maxorder_nonmon=3, threshold_sw=0.1, threshold_prec=0.1)
    return X, kw, call


def adapt_cross_case():
    """Seeded input of the cross-term structure search (adaptation_cross_terms, tm.py:4575-4950)."""
    rng = np.random.default_rng(32)
    n = 300
    x0 = rng.standard_normal(n)
    x1 = 0.6 * x0 ** 2 + 0.6 * rng.standard_normal(n)
    X = np.column_stack((x0, x1))
    kw = dict(monotone=None, nonmonotone=None, monotonicity='integrated rectifier', adaptation=True,
              adaptation_map_type='cross-terms', adaptation_max_order=3, adaptation_max_iterations=3,
              quadrature_input={'order': 15}, verbose=False)
    call = dict(increment=1e-6, chronicle=False)
    return X, kw, call


def fresh_kwargs(case):
    """Deep copy of the constructor kwargs (the reference mutates quadrature_input, tm.py:224)."""
    return copy.deepcopy(case['kwargs'])


def random_coeffs(n_non, n_mon, seed, scale=0.3):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n_non + n_mon) * scale
