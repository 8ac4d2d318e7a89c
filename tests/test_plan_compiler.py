"""Host-side term compiler (plan.py) against the oracle, on the CPU: the compiled tables, interpreted
with numpy, must reproduce the oracle's basis matrices for every parity workload."""

import numpy as np
import pytest

from cases import cases, fresh_kwargs
from harness import rel_err
from plan_interp import PlanInterp
from ttm_oracle import OracleMap
from ttt_b200.plan import ComponentPlan, resolve_family

CASES = cases()


@pytest.mark.parametrize('name', sorted(CASES))
def test_compiled_tables_reproduce_oracle_basis(name):
    case = CASES[name]
    kw = fresh_kwargs(case)
    om = OracleMap(X=case['X'].copy(), **kw)
    fam, polyfunc, polyder, _ = resolve_family(kw.get('polynomial_type', 'hermite function'))
    sep = kw['monotonicity'].startswith('sep')
    Dtot = case['X'].shape[1]
    for k in range(om.D):
        c = k + om.skip_dimensions
        plan = ComponentPlan(k, c, Dtot, fam, polyfunc, polyder, kw['monotone'][k], kw['nonmonotone'][k],
                             om.special_terms, kw.get('linearization'))
        it = PlanInterp(plan.iblob, plan.dblob)
        assert rel_err(it.terms(0, om.X), om.Psi_nonmon[k]) <= 1e-12
        assert rel_err(it.terms(1, om.X), om.Psi_mon[k]) <= 1e-12
        assert rel_err(it.nonmon_structured(om.X), om.Psi_nonmon[k]) <= 1e-12
        assert rel_err(it.mon_structured(om.X), om.Psi_mon[k]) <= 1e-12
        if sep:
            assert rel_err(it.terms(2, om.X), om.der_Psi_mon[k]) <= 1e-12
            assert np.array_equal(plan.lb, om.optimization_constraints_lb[k])
            assert np.array_equal(plan.ub, om.optimization_constraints_ub[k])
        assert plan.m_mon == len(om.coeffs_mon[k]) and plan.m_non == len(om.coeffs_nonmon[k])
        # alignment contract of the vector-loaded records
        h = plan.iblob
        from ttt_b200 import plan as PL
        assert h[PL.H_FAC_I] % 4 == 0 and h[PL.H_ENT_I] % 4 == 0 and h[PL.H_VAR_IDX] % 2 == 0
        assert h[PL.H_D_FAC] % 4 == 0 and h[PL.H_D_ENT] % 4 == 0 and h[PL.H_DENSE_VAR] % 4 == 0


def test_bad_options_raise_like_the_reference():
    with pytest.raises(Exception, match='Polynomial type not understood'):
        resolve_family('fourier')
    from ttt_b200.plan import parse_entry
    with pytest.raises(ValueError, match='not understood'):
        parse_entry('XBF 0', np.zeros(2, dtype=int), 0, True, None)
    with pytest.raises(Exception, match="'LIN' modifier"):
        parse_entry([0, 'LIN'], np.zeros(2, dtype=int), 0, True, None)


@pytest.mark.parametrize('name', [n for n in sorted(CASES) if any(
    type(e) == str for k in range(len(CASES[n]['kwargs']['monotone']))
    for e in CASES[n]['kwargs']['monotone'][k] + CASES[n]['kwargs']['nonmonotone'][k])])
def test_refresh_special_equals_rebuild_and_tables_do_not_depend_on_the_data(name):
    """reset() moves the special-term centres / scales (tm.py:800): patching them into the double blob must equal a
    full recompilation, and the int blob (factor table, term lists) must not change -- also when centres COINCIDE
    (tied quantiles of a discrete column, ST_scale_mode='static'), which used to merge factors."""
    import copy
    case = CASES[name]
    kw = fresh_kwargs(case)
    om = OracleMap(X=case['X'].copy(), **kw)
    fam, polyfunc, polyder, _ = resolve_family(kw.get('polynomial_type', 'hermite function'))
    Dtot = case['X'].shape[1]
    rng = np.random.default_rng(0)
    for k in range(om.D):
        c = k + om.skip_dimensions
        args = (k, c, Dtot, fam, polyfunc, polyder, kw['monotone'][k], kw['nonmonotone'][k])
        plan = ComponentPlan(*args, om.special_terms, kw.get('linearization'))
        for variant in ('moved', 'tied'):
            st = copy.deepcopy(om.special_terms)

            def perturb(d):
                for key, v in d.items():
                    if key == 'cross-terms':
                        perturb(v)
                    elif len(v['centers']):
                        if variant == 'moved':
                            v['centers'] = v['centers'] + rng.standard_normal(len(v['centers'])) * 0.1
                            v['scales'] = v['scales'] * (1.0 + 0.1 * rng.random(len(v['scales'])))
                        else:
                            v['centers'] = np.zeros_like(v['centers'])
                            v['scales'] = np.ones_like(v['scales'])
            perturb(st[c])
            fresh = ComponentPlan(*args, st, kw.get('linearization'))
            assert np.array_equal(fresh.iblob, plan.iblob), (name, k, variant)
            plan.refresh_special(st)
            assert np.array_equal(plan.dblob, fresh.dblob), (name, k, variant)


def test_donor_map_equals_the_all_pairs_definition():
    """plan.donor_map (one pass over the roots) against the definition it replaces: the longest special-term-free list
    of which the component's list is a prefix; first index among equally long ones."""
    from cases import c4_terms, c5_terms
    from ttt_b200.plan import donor_map

    def brute(non):
        out = {}
        for k, spec in enumerate(non):
            best = k
            if not any(type(e) == str for e in spec):
                for kk, other in enumerate(non):
                    if len(other) > len(non[best]) and other[:len(spec)] == spec and \
                            not any(type(e) == str for e in other):
                        best = kk
            out[k] = best
        return out

    rng = np.random.default_rng(3)
    lists = [c4_terms(12)[1], c5_terms(9)[1]]
    # two interleaved families of prefixes, a list with a special term, and an unrelated list
    fam_a = [[[]] + [[j] for j in range(n)] for n in (0, 2, 5, 3)]
    fam_b = [[[]] + [[j, j] for j in range(n)] for n in (1, 4, 2)]
    mixed = fam_a + fam_b + [[[], 'iRBF 0', [0]], [[1, 1, 1]]]
    lists.append([mixed[i] for i in rng.permutation(len(mixed))])
    for non in lists:
        got, want = donor_map(non), brute(non)
        for k in range(len(non)):
            # identical lists at different positions may share a donor where the all-pairs search kept k itself
            assert got[k] == want[k] or (non[got[k]][:len(non[k])] == non[k] and len(non[got[k]]) >= len(non[want[k]])), k
            assert non[got[k]][:len(non[k])] == non[k]
