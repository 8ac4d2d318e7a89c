"""Map adaptation on the CUDA path (adapt_map tm.py:373-657, adaptation_cross_terms :4575-4950) against fixtures
produced by the unmodified reference on the same seeded ensembles (tests/golden/make_golden_adapt.py): the searches
must select the same terms; the fitted maps must agree to the optimizer tolerance."""

import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

from cases import adapt_separable_case, adapt_cross_case
from harness import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def plain(spec):
    """Term lists with numpy integers turned into ints (the reference builds some entries from np.arange)."""
    return [[e if isinstance(e, str) else [x if isinstance(x, str) else int(x) for x in e] for e in comp] for comp in spec]


def lists_of(gold, key):
    return plain(eval(str(gold[key]), {'np': np}))


def test_separable_adaptation_selects_the_reference_terms():
    from transport_map import transport_map
    gold = np.load(os.path.join(GOLD, 'adapt_separable.npz'))
    X, kw, call = adapt_separable_case()
    tm = transport_map(X=X.copy(), **kw)
    assert tm.D == 3 and tm.monotone == [[[]]] * 3                   # dummy map of the adaptation ctor, tm.py:331-345
    with redirect_stdout(io.StringIO()):
        tm.adapt_map(**call)
    assert plain(tm.monotone) == lists_of(gold, 'monotone')
    assert plain(tm.nonmonotone) == lists_of(gold, 'nonmonotone')
    assert np.array_equal(tm.maporders, gold['maporders'])
    for k in range(tm.D):
        assert rel_err(tm.coeffs_mon[k], gold['coeffs_mon_%d' % k]) <= 1e-6, k
        assert rel_err(tm.coeffs_nonmon[k], gold['coeffs_nonmon_%d' % k]) <= 1e-6, k
    assert rel_err(tm.map(), gold['map_train']) <= 1e-6


def test_cross_term_adaptation_selects_the_reference_terms():
    """The re-fits use scipy's finite-difference gradient of the objective (the reference passes no jac, :4893): the
    selected multi-indices must be identical; coefficients agree to the finite-difference noise."""
    from transport_map import transport_map
    gold = np.load(os.path.join(GOLD, 'adapt_cross_terms.npz'))
    X, kw, call = adapt_cross_case()
    tm = transport_map(X=X.copy(), **kw)
    with redirect_stdout(io.StringIO()):
        tm.adaptation_cross_terms(**call)
    assert plain(tm.monotone) == lists_of(gold, 'monotone')
    assert plain(tm.nonmonotone) == lists_of(gold, 'nonmonotone')
    assert np.array_equal(tm.multi_index_matrix, gold['multi_index_matrix'])
    for k in range(tm.D):
        assert rel_err(tm.coeffs_mon[k], gold['coeffs_mon_%d' % k]) <= 1e-4, k
        assert rel_err(tm.coeffs_nonmon[k], gold['coeffs_nonmon_%d' % k]) <= 1e-4, k
    assert rel_err(tm.map(X.copy()), gold['map_train']) <= 1e-4


def test_single_component_recompile_keeps_the_other_components():
    """function_constructor_alternative(k) (tm.py:1263, partial construction): only component k's tables change."""
    from transport_map import transport_map
    from cases import synthetic_samples, c4_terms
    X = synthetic_samples(500, 3, seed=12)
    mon, non = c4_terms(3)
    tm = transport_map(X=X.copy(), monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                       quadrature_input={'order': 12}, verbose=False)
    rng = np.random.default_rng(0)
    c2 = rng.standard_normal(len(non[2]) + len(mon[2])) * 0.1
    before = tm.objective_function(c2, 2, len(non[2]))
    tm.nonmonotone[1] = [[], [0], [0, 0, 'HF']]
    tm.function_constructor_alternative(k=1)
    assert tm.objective_function(c2, 2, len(non[2])) == before
    c1 = rng.standard_normal(3 + len(mon[1])) * 0.1
    fresh = transport_map(X=X.copy(), monotone=mon, nonmonotone=[non[0], [[], [0], [0, 0, 'HF']], non[2]],
                          monotonicity='integrated rectifier', quadrature_input={'order': 12}, verbose=False)
    assert tm.objective_function(c1, 1, 3) == fresh.objective_function(c1, 1, 3)
    assert np.array_equal(tm.objective_function_jacobian(c1, 1, 3), fresh.objective_function_jacobian(c1, 1, 3))
