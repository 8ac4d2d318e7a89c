"""
Implementation-agnostic driver for the parity workloads in tests/cases.py.

`run_case(make, case)` exercises one map through the reference's public API
(constructor, Psi_* attributes, objective_function / _jacobian, optimize, map,
inverse_map, density evaluators) and returns {name: ndarray}.  `make(X, **kwargs)`
builds the object: the unmodified reference (tests/golden/make_golden.py), the
CPU oracle, or the CUDA-backed drop-in class.
"""

import copy

import numpy as np

from cases import fresh_kwargs, random_coeffs


def _log_gauss(X):
    return -0.5 * np.sum(X ** 2, axis=1) - 0.5 * X.shape[1] * np.log(2 * np.pi)


def run_case(make, case, fitted=None):
    """fitted: optional dict with 'coeffs_mon_k'/'coeffs_nonmon_k' to impose instead of optimising
    (lets the GPU tests check map/inverse on *identical* coefficients)."""
    out = {}
    kw = fresh_kwargs(case)
    tm = make(copy.copy(case['X']), **kw)
    sep = kw['monotonicity'].lower().startswith('sep')
    if 'reset_X' in case:
        tm.reset(copy.copy(case['reset_X']))
    D = tm.D
    out['X_standardized'] = np.array(tm.X)
    if kw.get('standardize_samples', True):
        out['X_mean'] = np.array(tm.X_mean)
        out['X_std'] = np.array(tm.X_std)
    for k in range(D):
        out['Psi_mon_%d' % k] = np.array(tm.Psi_mon[k])
        out['Psi_nonmon_%d' % k] = np.array(tm.Psi_nonmon[k])
        if sep:
            out['der_Psi_mon_%d' % k] = np.array(tm.der_Psi_mon[k])

    # objective / gradient at random coefficients (integrated rectifier only)
    if not sep and case.get('objgrad', True):
        for k in range(D):
            div = len(tm.coeffs_nonmon[k])
            c = random_coeffs(div, len(tm.coeffs_mon[k]), seed=100 + k)
            out['obj_%d' % k] = np.asarray(tm.objective_function(c.copy(), k, div))
            out['grad_%d' % k] = np.asarray(tm.objective_function_jacobian(c.copy(), k, div))

    # coefficients: fitted (or imposed), else random
    if fitted is not None:
        for k in range(D):
            tm.coeffs_mon[k] = np.array(fitted['coeffs_mon_%d' % k])
            tm.coeffs_nonmon[k] = np.array(fitted['coeffs_nonmon_%d' % k])
    elif case.get('fit', False):
        tm.optimize()
    else:
        for k in range(D):
            div = len(tm.coeffs_nonmon[k])
            c = random_coeffs(div, len(tm.coeffs_mon[k]), seed=200 + k, scale=0.2)
            tm.coeffs_nonmon[k] = c[:div].copy()
            tm.coeffs_mon[k] = np.abs(c[div:]) + 0.05 if sep else c[div:].copy()
    for k in range(D):
        out['coeffs_mon_%d' % k] = np.array(tm.coeffs_mon[k])
        out['coeffs_nonmon_%d' % k] = np.array(tm.coeffs_nonmon[k])

    Dtot = case['X'].shape[1]
    skip = Dtot - D
    rng = np.random.default_rng(1234)
    mu, sd = case.get('reset_X', case['X']).mean(axis=0), case.get('reset_X', case['X']).std(axis=0)
    Xe = mu + sd * rng.standard_normal((96, Dtot))
    out['map_X'] = tm.map(copy.copy(Xe))
    out['map_train'] = tm.map()

    n = case.get('n_inverse', 0)
    if n:
        Z = rng.standard_normal((n, D))
        xs = Xe[:1].repeat(n, axis=0)[:, :skip] + 0.1 * rng.standard_normal((n, skip)) if skip else None
        if 'ystar' in case:
            Z = tm.map(copy.copy(case['reset_X']))[:n]
            xs = np.full((n, skip), case['ystar'])
        modes = [True, False] if sep else [True]
        for alt in modes:
            tm.alternate_root_finding = alt
            tag = 'table' if (alt and sep) else 'bisect'
            out['inverse_%s' % tag] = tm.inverse_map(copy.copy(Z), None if xs is None else copy.copy(xs))
            if 'cond' in case:                      # full map + X_star: branch C of inverse_map
                E = case['cond']
                out['inverse_cond_%s' % tag] = tm.inverse_map(copy.copy(Z[:, E:]), copy.copy(Xe[:n, :E]) if n <= 96
                                                            else copy.copy(np.resize(Xe[:, :E], (n, E))))
        tm.alternate_root_finding = True

    if case.get('densities', False):
        if skip:
            out['pullback'] = tm.evaluate_pullback_density(copy.copy(Xe[:, skip:]), X_star=copy.copy(Xe[:, :skip]))
            Zd = rng.standard_normal((96, D))
            out['pushforward'] = tm.evaluate_pushforward_density(copy.copy(Zd), _log_gauss, X_star=copy.copy(Xe[:, :skip]))
        else:
            out['pullback'] = tm.evaluate_pullback_density(copy.copy(Xe))
            Zd = rng.standard_normal((96, D))
            out['pushforward'] = tm.evaluate_pushforward_density(copy.copy(Zd), _log_gauss)
    return out


def rel_err(a, b):
    """max |a-b| / max(1, |b|)  (the SURVEY.md 7.2 parity measure)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.shape != b.shape:
        return np.inf
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))
