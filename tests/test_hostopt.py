"""hostopt.quasi_newton (the O(n^2) BFGS driver of the integrated-rectifier fits) against scipy's BFGS, which the
reference calls (tm.py:3252-3257): same line search, same constants -- the iterates must coincide up to rounding."""

import importlib.util
import os

import numpy as np
import pytest
from scipy.optimize import minimize

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    'ttt_hostopt', os.path.join(HERE, '..', 'triangular-transport-toolbox_b200', 'hostopt.py'))
H = importlib.util.module_from_spec(spec)
spec.loader.exec_module(H)


def problems():
    rng = np.random.default_rng(0)
    out = []
    for n in (3, 20, 120):
        A = rng.standard_normal((n, n))
        A = A @ A.T / n + 0.1 * np.eye(n)
        b = rng.standard_normal(n)
        W = rng.standard_normal((3 * n, n)) / np.sqrt(n)
        # strictly convex, transport-map-like: quadratic + sum of -log(softplus-ish positive affine) terms
        def f(x, A=A, b=b, W=W):
            z = W @ x
            return 0.5 * x @ A @ x - b @ x + np.sum(np.logaddexp(0.0, z))
        def g(x, A=A, b=b, W=W):
            z = W @ x
            return A @ x - b + W.T @ (1.0 / (1.0 + np.exp(-z)))
        out.append((n, f, g, rng.standard_normal(n) * 0.1))
    return out


@pytest.mark.parametrize('case', problems(), ids=lambda c: 'n%d' % c[0])
def test_matches_scipy_bfgs(case):
    assert H.available()
    n, f, g, x0 = case
    ref = minimize(f, x0, jac=g, method='BFGS')
    res = H.quasi_newton(lambda x: (f(x), g(x)), x0)
    assert res.success and ref.success
    assert res.nit == ref.nit                                     # same trajectory, not just the same minimum
    assert np.max(np.abs(res.x - ref.x)) <= 1e-7                  # rounding of the rank-two update, amplified over ~30 steps
    assert abs(res.fun - ref.fun) <= 1e-12 * max(1.0, abs(ref.fun))
    assert res.nfev <= ref.nfev                                   # fused (f, grad): never more evaluations than scipy


def test_reports_line_search_failure_like_scipy():
    f = lambda x: float(np.sum(np.abs(x)) + 1e-3 * np.sum(x ** 2))   # kink at the minimum: the Wolfe search gives up
    g = lambda x: np.sign(x) + 2e-3 * x
    x0 = np.array([0.3, -0.2])
    ref = minimize(f, x0, jac=g, method='BFGS')
    res = H.quasi_newton(lambda x: (f(x), g(x)), x0)
    assert res.status == ref.status


def test_lbfgsb_lockstep_reproduces_scipy_per_problem():
    """Several bounded problems of different size and iteration count advanced together: every one must end on
    scipy.optimize.minimize(method='L-BFGS-B')'s iterate, bit for bit, with the same nit / nfev / status."""
    from scipy.optimize import minimize
    hostopt = H
    assert hostopt.lbfgsb_available()
    rng = np.random.default_rng(5)
    probs = []
    for n in (1, 3, 6, 11):
        Q = rng.standard_normal((n, n))
        Q = Q @ Q.T / n + 0.1 * np.eye(n)
        c = rng.standard_normal(n)

        def fg(x, Q=Q, c=c):                       # the separable fits' shape: quadratic minus log of a positive sum
            s = 1e-3 + np.sum(x) + 0.5
            return 0.5 * x @ Q @ x + c @ x - np.log(s), Q @ x + c - 1.0 / s
        lb = np.zeros(n)
        ub = np.full(n, np.inf)
        if n > 2:
            ub[1] = 0.25
        probs.append((fg, rng.uniform(0.1, 1.0, n), lb, ub))
    box = {}
    res = hostopt.lbfgsb_lockstep([p[1] for p in probs], [(p[2], p[3]) for p in probs],
                                  lambda i, x: box.__setitem__(i, probs[i][0](x.copy())), lambda i: box.pop(i))
    for (fg, x0, lb, ub), mine in zip(probs, res):
        ref = minimize(fg, x0, jac=True, method='L-BFGS-B',
                       bounds=[(l, None if np.isinf(u) else u) for l, u in zip(lb, ub)])
        assert np.array_equal(ref.x, mine.x)
        assert (ref.nit, ref.nfev, ref.status, ref.fun) == (mine.nit, mine.nfev, mine.status, mine.fun)
        assert ref.message == mine.message
