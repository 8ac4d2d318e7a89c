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


@pytest.mark.parametrize('D,E,mixed', [(20, 3, False), (21, 9, True), (150, 5, False)])
def test_packed_operands_of_the_fused_kernels_reproduce_the_nonmonotone_sums(D, E, mixed):
    """plan.pack_fused_operands (Apack of K-inv-fused, Rpack of K-inv-rect / K-map-rect): contracting the packed
    coefficient*scale with the features the kernels form ({He1, He2 e, He3 e} or the 6-slot set) must reproduce the
    oracle's nonmonotone part  Psi_non a  of every component -- block layout, triangular zero padding, slot order,
    Hermite-function scales and the split between the conditioning block (Rpack) and the solved columns."""
    from cases import c5_terms, synthetic_samples, headline_sep_coeffs
    from ttt_b200 import plan as PL
    mon, non = c5_terms(D)
    if mixed:
        non = [[[]] + [t for j in range(k) for t in ([j], [j, 'HF'], [j, j], [j, j, 'HF'], [j, j, j], [j, j, j, 'HF'])]
               for k in range(D)]
    X = synthetic_samples(300, D, seed=31)
    kw = dict(monotone=mon, nonmonotone=non, monotonicity='separable monotonicity')
    om = OracleMap(X=X.copy(), **kw)
    fam, polyfunc, polyder, _ = resolve_family('hermite function')
    plans = [ComponentPlan(k, k, D, fam, polyfunc, polyder, mon[k], non[k], om.special_terms, None) for k in range(D)]
    cm, cn = headline_sep_coeffs(mon, non)
    Xs = om.X                                                   # standardised samples, as the kernels see them
    n = Xs.shape[0]

    def feats(x, ns):
        ga = np.exp(-0.25 * x * x)
        P2, P3 = x * x - 1.0, x * (x * x - 3.0)
        return [x, P2 * ga, P3 * ga] if ns == 3 else [x, x * ga, P2, P2 * ga, P3, P3 * ga]

    # --- conditional inverse: components E .. D-1, split at E
    sub = plans[E:]
    st = PL.pack_fused_operands(sub, E, E)
    assert st is not None and st['ns'] == (6 if mixed else 3)
    ns, CB, ncomp = st['ns'], PL.FUSED_CB, D - E
    cat = np.concatenate([cn[k] for k in range(E, D)])
    val = cat[st['src']] * st['sc']
    A = np.zeros(PL.fused_apack_doubles(ncomp, E, ns))
    A[st['dst']] = val
    R = np.zeros(PL.rect_rpack_doubles(ncomp, E, ns))
    R[st['rdst']] = val[st['rkeep']]
    rp = (E + 7) // 8 * 8
    for j in (0, 1, ncomp // 2, ncomp - 1):
        k = E + j
        want = om.Psi_nonmon[k] @ cn[k]
        a0 = sum(cat[q] for q in st['const_src'][st['const_ptr'][j]:st['const_ptr'][j + 1]])
        b, jj = divmod(j, CB)
        row0 = b * (E + CB) + CB * b * (b - 1) // 2
        walk_all = np.full(n, a0)
        for v in range(min(D, E + CB * b + CB)):               # the rows block b holds (zero beyond the predecessors)
            f = feats(Xs[:, v], ns)
            for q in range(ns):
                walk_all += f[q] * A[((row0 + v) * CB + jj) * ns + q]
        assert rel_err(walk_all, want) <= 1e-12, (j, 'Apack')
        # split: Rpack carries the columns < E, the walk (same Apack rows) the columns >= E
        base = np.zeros(n)
        for v in range(E):
            f = feats(Xs[:, v], ns)
            for q in range(ns):
                base += f[q] * R[(((j // 128) * rp + v) * ns + q) * 128 + j % 128]
        walk = np.full(n, a0)
        for v in range(E, min(D, E + CB * b + CB)):
            f = feats(Xs[:, v], ns)
            for q in range(ns):
                walk += f[q] * A[((row0 + v) * CB + jj) * ns + q]
        assert rel_err(base + walk, want) <= 1e-12, (j, 'Rpack + Apack')
    # --- forward map: all components, every predecessor column in Rpack
    gm = PL.pack_fused_operands(plans, 0, D - 1, want_apack=False)
    cat = np.concatenate(cn)
    Rm = np.zeros(PL.rect_rpack_doubles(D, D - 1, gm['ns']))
    Rm[gm['rdst']] = (cat[gm['src']] * gm['sc'])[gm['rkeep']]
    rp = (D - 1 + 7) // 8 * 8
    for k in (0, 1, D // 2, D - 1):
        a0 = sum(cat[q] for q in gm['const_src'][gm['const_ptr'][k]:gm['const_ptr'][k + 1]])
        got = np.full(n, a0)
        for v in range(D - 1):
            f = feats(Xs[:, v], gm['ns'])
            for q in range(gm['ns']):
                got += f[q] * Rm[(((k // 128) * rp + v) * gm['ns'] + q) * 128 + k % 128]
        assert rel_err(got, om.Psi_nonmon[k] @ cn[k]) <= 1e-12, (k, 'map Rpack')
