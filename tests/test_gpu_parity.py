"""GPU parity tests proper: the CUDA path, called through the drop-in class / C ABI, against
(a) the golden fixtures produced by the unmodified reference and (b) the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): basis matrices, objective and gradient <= 1e-10 relative;
fitted coefficients and map outputs <= 1e-6 relative under the same optimizer; inverse samples to the
reference root-finding tolerance (bisection stops on |S(x) - z| <= 1e-9; the table inverse is a
deterministic interpolation: <= 1e-10)."""

import os

import numpy as np
import pytest

from cases import cases, fresh_kwargs, synthetic_samples, c4_terms, c5_terms, ex01_terms
from harness import run_case, rel_err

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = cases()

TOL_BASIS = 1e-10     # Psi, J, grad
TOL_MAP = 1e-9        # forward map with identical coefficients (1e-6 required)
TOL_FIT = 1e-6        # coefficients after optimize()
TOL_BISECT = 2e-8     # both sides stop on a 1e-9 residual; slopes are O(0.1..10)
TOL_TABLE = 1e-10


def make_cuda(X, **kw):
    from transport_map import transport_map
    return transport_map(X=X, **kw)


def tol_for(key):
    if key.startswith(('Psi', 'der_Psi', 'obj_', 'grad_', 'X_')):
        return TOL_BASIS
    if key.startswith('inverse') and key.endswith('bisect'):
        return TOL_BISECT
    if key.startswith('inverse'):
        return TOL_TABLE
    if key.startswith('coeffs'):
        return 0.0      # imposed
    if key in ('pullback', 'pushforward'):
        return 1e-9
    return TOL_MAP


@pytest.mark.parametrize('name', sorted(CASES))
def test_cuda_matches_reference_fixture(name):
    """Everything except the optimizer: coefficients are imposed from the fixture so that map / inverse /
    densities are compared on identical inputs."""
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    res = run_case(make_cuda, CASES[name], fitted=gold)
    assert set(res) == set(gold.files) - {'_versions'}
    for key, val in res.items():
        err = rel_err(val, gold[key])
        assert err <= tol_for(key), (name, key, err)


FIT_CASES = [n for n in sorted(CASES) if CASES[n].get('fit')]


@pytest.mark.parametrize('name', FIT_CASES)
def test_cuda_fit_matches_reference_coefficients(name):
    """optimize() end to end (scipy BFGS / L-BFGS-B on the host fed by the CUDA objective)."""
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    case = CASES[name]
    tm = make_cuda(case['X'].copy(), **fresh_kwargs(case))
    if 'reset_X' in case:
        tm.reset(case['reset_X'].copy())
    tm.optimize()
    for k in range(tm.D):
        assert rel_err(tm.coeffs_mon[k], gold['coeffs_mon_%d' % k]) <= TOL_FIT, (name, k, 'mon')
        assert rel_err(tm.coeffs_nonmon[k], gold['coeffs_nonmon_%d' % k]) <= TOL_FIT, (name, k, 'nonmon')
    assert rel_err(tm.map(), gold['map_train']) <= TOL_FIT


def test_example01_known_answer():
    """The reference's shipped Example-01/02 coefficients: J_0 = 0.22517858233600668,
    J_1 = -0.7978830242276352, stationary to |grad| < 3e-5."""
    ka = np.load(os.path.join(GOLD, 'ex01_known_answer.npz'))
    mon, non = ex01_terms(10)
    tm = make_cuda(ka['X'].copy(), monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                   quadrature_input={'order': 25}, verbose=False)
    expected = {0: 0.22517858233600668, 1: -0.7978830242276352}
    for k in range(2):
        c, div = ka['full_coeffs_%d' % k], int(ka['full_div_%d' % k])
        J = tm.objective_function(c.copy(), k, div)
        g = tm.objective_function_jacobian(c.copy(), k, div)
        assert abs(J - expected[k]) <= 1e-10 * max(1, abs(expected[k]))
        assert rel_err(g, ka['full_grad_%d' % k]) <= TOL_BASIS
        assert np.linalg.norm(g) < 3e-5
        tm.coeffs_nonmon[k], tm.coeffs_mon[k] = c[:div].copy(), c[div:].copy()
    Z = tm.map(ka['X'].copy())
    assert rel_err(Z[:512], ka['full_map_head']) <= TOL_MAP
    assert rel_err(Z.std(axis=0), ka['full_map_std']) <= TOL_MAP
    # partial (conditional) map of Example 02
    tm2 = make_cuda(ka['X'].copy(), monotone=mon[1:], nonmonotone=non[1:], monotonicity='integrated rectifier',
                    quadrature_input={'order': 25}, verbose=False)
    c, div = ka['partial_coeffs_0'], int(ka['partial_div_0'])
    assert abs(tm2.objective_function(c.copy(), 0, div) - float(ka['partial_J_0'])) <= 1e-10
    assert rel_err(tm2.objective_function_jacobian(c.copy(), 0, div), ka['partial_grad_0']) <= TOL_BASIS


def test_against_live_oracle_medium_size():
    """Same seeded inputs, oracle evaluated here (sizes it finishes in seconds)."""
    from ttm_oracle import OracleMap
    X = synthetic_samples(20000, 6, seed=21)
    mon, non = c4_terms(6)
    kw = dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier', verbose=False)
    tm = make_cuda(X.copy(), quadrature_input={'order': 30}, **kw)
    om = OracleMap(X=X.copy(), quadrature_input={'order': 30}, **kw)
    rng = np.random.default_rng(5)
    for k in range(6):
        div = len(tm.coeffs_nonmon[k])
        c = rng.standard_normal(div + len(tm.coeffs_mon[k])) * 0.1
        assert abs(tm.objective_function(c, k, div) - om.objective_function(c, k, div)) <= 1e-10
        assert rel_err(tm.objective_function_jacobian(c, k, div), om.objective_function_jacobian(c, k, div)) <= 1e-10


def test_full_size_properties_c4():
    """BASELINE config C4 at full size (N = 1M, D = 64, Q = 100 / 25): size-independent properties.
    (i)  J and grad are sample means: the value on the whole ensemble equals the weighted mean over two
         disjoint halves (linearity; standardisation disabled so both see the same columns);
    (ii) bit-reproducibility of repeated launches (fixed-order reductions);
    (iii) inverse_map(map(X)) round trip on a slice to the bisection tolerance."""
    N, D = 1_000_000, 64
    X = synthetic_samples(N, D, seed=0)
    X = (X - X.mean(axis=0)) / X.std(axis=0)
    mon, non = c4_terms(D)
    kw = dict(monotone=mon, nonmonotone=non, monotonicity='integrated rectifier', verbose=False,
              standardize_samples=False)
    tm = make_cuda(X, quadrature_input={'order': 25}, **kw)
    n1 = 400_000
    ta = make_cuda(X[:n1], quadrature_input={'order': 25}, **kw)
    tb = make_cuda(X[n1:], quadrature_input={'order': 25}, **kw)
    rng = np.random.default_rng(0)
    for k in (0, 31, 63):
        div = len(tm.coeffs_nonmon[k])
        c = rng.standard_normal(div + len(tm.coeffs_mon[k])) * 0.05
        J, g = tm.objective_function(c, k, div), tm.objective_function_jacobian(c, k, div)
        J2, g2 = tm.objective_function(c + 0.0, k, div), tm.objective_function_jacobian(c + 0.0, k, div)
        assert J == J2 and np.array_equal(g, g2)
        Ja, ga = ta.objective_function(c, k, div), ta.objective_function_jacobian(c, k, div)
        Jb, gb = tb.objective_function(c, k, div), tb.objective_function_jacobian(c, k, div)
        assert abs(J - (n1 * Ja + (N - n1) * Jb) / N) <= 1e-12 * max(1, abs(J))
        assert rel_err(g, (n1 * ga + (N - n1) * gb) / N) <= 1e-12
    for k in range(4):
        div = len(tm.coeffs_nonmon[k])
        c = rng.standard_normal(div + len(tm.coeffs_mon[k])) * 0.05
        ta.coeffs_nonmon[k], ta.coeffs_mon[k] = c[:div].copy(), c[div:].copy()
    # round trip on the first 4 components of a partial copy
    t4 = make_cuda(X[:50_000, :4], monotone=mon[:4], nonmonotone=non[:4], monotonicity='integrated rectifier',
                   verbose=False, quadrature_input={'order': 25})
    for k in range(4):
        t4.coeffs_nonmon[k], t4.coeffs_mon[k] = ta.coeffs_nonmon[k], ta.coeffs_mon[k]
    Z = t4.map(X[:50_000, :4].copy())
    Xr = t4.inverse_map(Z)
    assert np.max(np.abs(t4.map(Xr) - Z)) <= 5e-9


def test_full_size_properties_c5_inverse():
    """BASELINE config C5 shape (separable LET/iRBF/RET map, conditional inverse): round trip
    map(inverse_map(Z, X*)) == Z through both root finders, and table vs bisection agreement."""
    D, E, Ntrain, Ns = 24, 12, 4000, 200_000
    X = synthetic_samples(Ntrain, D, seed=3)
    mon, non = c5_terms(D)
    tm = make_cuda(X.copy(), monotone=mon, nonmonotone=non, monotonicity='separable monotonicity', verbose=False)
    tm.optimize()
    rng = np.random.default_rng(1)
    Xnew = synthetic_samples(Ns, D, seed=4)
    Z = rng.standard_normal((Ns, D - E))
    tm.alternate_root_finding = False
    Xb = tm.inverse_map(Z, X_star=Xnew[:, :E].copy())
    assert Xb.shape == (Ns, D)
    Zb = tm.map(Xb)[:, E:]
    assert np.max(np.abs(Zb - Z)) <= 5e-9
    assert np.max(np.abs(Xb[:, :E] - Xnew[:, :E])) <= 1e-12
    tm.alternate_root_finding = True
    Xt = tm.inverse_map(Z, X_star=Xnew[:, :E].copy())
    inside = np.all(np.abs(Z) < 3, axis=1)
    assert np.max(np.abs(Xt[inside] - Xb[inside])) <= 1e-3      # table resolution 0.02 with linear interpolation


def test_errors_match_reference_messages():
    X = synthetic_samples(64, 2, seed=1)
    mon, non = c4_terms(2)
    with pytest.raises(ValueError, match="'ST_scale_mode' must be either 'dynamic' or 'static'."):
        make_cuda(X, monotone=mon, nonmonotone=non, ST_scale_mode='wide')
    with pytest.raises(ValueError, match="not understood"):
        make_cuda(X, monotone=mon, nonmonotone=non, monotonicity='convex')
    with pytest.raises(Exception, match="Polynomial type not understood"):
        make_cuda(X, monotone=mon, nonmonotone=non, polynomial_type='fourier')
    with pytest.raises(ValueError, match="'standardization' must be either 'standard' or 'quantiles'."):
        make_cuda(X, monotone=mon, nonmonotone=non, standardization='zscore')
    tm = make_cuda(X, monotone=mon, nonmonotone=non, verbose=False)
    with pytest.raises(AssertionError):
        tm.evaluate_pullback_density(X)
    with pytest.raises(Exception, match='two-dimensional'):
        tm.reset(X[:, 0])
