"""Pin the CPU oracle to the reference: every golden fixture (outputs of the unmodified
reference, tests/golden/make_golden.py) and the reference's shipped Example-01/02 known answers."""

import os

import numpy as np
import pytest

from cases import cases, ex01_terms
from harness import run_case, rel_err
from ttm_oracle import OracleMap

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = cases()


def make_oracle(X, **kw):
    return OracleMap(X=X, **kw)


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_reference_fixture(name):
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    res = run_case(make_oracle, CASES[name])
    assert set(res) == set(gold.files) - {'_versions'}
    for key, val in res.items():
        # the restatement performs the same numpy calls in the same order: 1e-12 is generous
        assert rel_err(val, gold[key]) <= 1e-12, (name, key, rel_err(val, gold[key]))


def test_oracle_reproduces_shipped_known_answers():
    """Example 01/02 pickled coefficients are a stationary point with the published J values
    (SURVEY.md section 4): J_0 = 0.22517858233600668, J_1 = -0.7978830242276352."""
    ka = np.load(os.path.join(GOLD, 'ex01_known_answer.npz'))
    mon, non = ex01_terms(10)
    om = OracleMap(X=ka['X'].copy(), monotone=mon, nonmonotone=non, monotonicity='integrated rectifier',
                   quadrature_input={'order': 25})
    expected = {0: 0.22517858233600668, 1: -0.7978830242276352}
    for k in range(2):
        c, div = ka['full_coeffs_%d' % k], int(ka['full_div_%d' % k])
        J = om.objective_function(c.copy(), k, div)
        g = om.objective_function_jacobian(c.copy(), k, div)
        assert abs(J - expected[k]) <= 1e-13
        assert abs(J - float(ka['full_J_%d' % k])) <= 1e-13
        assert np.max(np.abs(g - ka['full_grad_%d' % k])) <= 1e-12
        assert np.linalg.norm(g) < 3e-5
