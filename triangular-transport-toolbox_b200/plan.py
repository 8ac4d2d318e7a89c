"""
Host-side term-table compiler.

The reference turns every entry of `monotone[k]` / `nonmonotone[k]` into Python source and `exec`s
it (write_basis_function tm.py:823-1261, function_constructor_alternative :1263-1856,
function_derivative_constructor_alternative :1860-2134).  Here the same grammar is compiled into
flat tables that the CUDA kernels interpret (layout: csrc/ttm_common.cuh):

    factor  = (column, kind, order | scale, scale2, mu, sigma)      deduplicated per component
    term    = product of factors                                     CSR lists `non`, `mon`, `dmon`

plus the structures the fused integrated-rectifier kernel needs (nonmonotone terms grouped by
variable; monotone terms split into an outer product over x_<c and one *slot* of x_c).

Only constants are computed here (Hermite-function normalisers, centre/scale look-ups); all sample
arithmetic happens on the device.  "tm.py" = /root/reference/transport_map.py.
"""

import itertools

import numpy as np

# factor kinds / families / header indices: keep in sync with csrc/ttm_common.cuh
F_POLY, F_POLY_HF, F_RBF, F_IRBF, F_LET, F_RET = 0, 1, 2, 3, 4, 5
F_DPOLY, F_DPOLY_HF, F_DRBF, F_DIRBF, F_DLET, F_DRET, F_ZERO, F_ONE = 6, 7, 8, 9, 10, 11, 12, 13
FAM_POWER, FAM_HERMITE, FAM_HERMITE_E, FAM_CHEBYSHEV, FAM_LAGUERRE, FAM_LEGENDRE = 0, 1, 2, 3, 4, 5
PLAN_MAGIC = 0x54544d31
(H_MAGIC, H_DTOT, H_C, H_FAMILY, H_NFAC, H_FAC_I,
 H_M_NON, H_NON_PTR, H_NON_FAC, H_M_MON, H_MON_PTR, H_MON_FAC, H_M_DMON, H_DMON_PTR, H_DMON_FAC,
 H_NCONST, H_CONST_IDX, H_NVARS, H_VAR_IDX, H_VAR_PTR, H_ENT_I, H_NMULTI, H_MULTI_IDX,
 H_MAXORD, H_HAS_PLAIN, H_HAS_HF, H_NST, H_NSLOT, H_SLOT_PTR, H_SLOT_TERM, H_OUT_PTR, H_OUT_FAC, H_ST_FAC,
 H_D_FAC, H_D_ENT, H_D_SLOT_SCALE, H_D_REC, H_NON_MAXVAR,
 H_NDENSE, H_DENSE_VAR, H_DENSE_IDX, H_DENSE_MAXORD, H_D_DENSE_SCALE, H_NACTIVE, H_NOUTFAC) = range(45)
H_SIZE = 48

_ST_KIND = {'rbf': F_RBF, 'irbf': F_IRBF, 'let': F_LET, 'ret': F_RET}
_ST_DKIND = {'rbf': F_DRBF, 'irbf': F_DIRBF, 'let': F_DLET, 'ret': F_DRET}


def resolve_family(polynomial_type):
    """polynomial_type keyword -> (family id, numpy class, numpy derivative fn, canonical name); tm.py:271-304."""
    P = np.polynomial
    t = polynomial_type.lower()
    if t in ('standard', 'polynomial', 'power series'):
        return FAM_POWER, P.polynomial.Polynomial, P.polynomial.polyder, polynomial_type
    if t in ('hermite', "phycisist's hermite", 'phycisists hermite'):
        return FAM_HERMITE, P.hermite.Hermite, P.hermite.hermder, polynomial_type
    if t in ('hermite_e', "probabilist's hermite", 'probabilists hermite'):
        return FAM_HERMITE_E, P.hermite_e.HermiteE, P.hermite_e.hermeder, polynomial_type
    if t == 'chebyshev':
        return FAM_CHEBYSHEV, P.chebyshev.Chebyshev, P.chebyshev.chebder, polynomial_type
    if t == 'laguerre':
        return FAM_LAGUERRE, P.laguerre.Laguerre, P.laguerre.lagder, polynomial_type
    if t == 'legendre':
        return FAM_LEGENDRE, P.legendre.Legendre, P.legendre.legder, polynomial_type
    if t in ('hermite function', 'hermite_function', 'hermite functions'):
        return FAM_HERMITE_E, P.hermite_e.HermiteE, P.hermite_e.hermeder, 'hermite function'
    raise Exception("Polynomial type not understood. The variable polynomial_type should be either "
                    "'power series', 'hermite', 'hermite_e', 'chebyshev', 'laguerre', or 'legendre'.")


_HF_CACHE = {}


def hf_normaliser(polyfunc, order):
    """1 / max |P_n(x) exp(-x^2/4)| on linspace(-100, 100, 100001): a *grid* maximum, tm.py:1102-1109.
    Cached per (family, order): the reference recomputes it for every factor (191 s at D=256)."""
    key = (polyfunc.__name__, int(order))
    if key not in _HF_CACHE:
        g = np.linspace(-100, 100, 100001)
        _HF_CACHE[key] = 1 / np.max(np.abs(polyfunc([0.] * order + [1.])(g) * np.exp(-g ** 2 / 4)))
    return _HF_CACHE[key]


def parse_entry(entry, st_counter, c, monotone_fn, linearization):
    """One spec entry -> record.  Case split of write_basis_function (tm.py:880-1090) and the special
    term bookkeeping of function_constructor_alternative (tm.py:1392-1427)."""
    if entry == []:
        return {'type': 'const'}
    if type(entry) == str:
        kind, i = entry.split(' ')
        i = int(i)
        if kind.lower() not in _ST_KIND:
            raise ValueError("Special term '" + str(kind) + "' not understood. Currently, only LET, RET, "
                             "iRBF, and RBF are implemented.")
        rec = {'type': 'st', 'kind': kind.lower(), 'var': i, 'cross': bool(monotone_fn and i != c),
               'slot': int(st_counter[i])}
        st_counter[i] += 1
        return rec
    hf = any(e == 'HF' for e in entry)
    lin = any(e == 'LIN' for e in entry)
    if lin and linearization is None:
        raise Exception("'LIN' modifier specified in variable monotone, but the variable linearization is "
                        "defined as None. Please specify a scalar linearization or remove the 'LIN' modifier.")
    # sorted distinct variables and their multiplicities (np.unique(..., return_counts=True) of tm.py:1430,
    # written out: the lists have a handful of entries and this runs once per term)
    counts = {}
    for e in entry:
        if type(e) != str:
            counts[int(e)] = counts.get(int(e), 0) + 1
    ui = sorted(counts)
    # NB 'LIN' is a numerical no-op in the reference: the factor is evaluated as is.
    return {'type': 'poly', 'hf': hf, 'vars': ui, 'orders': [counts[u] for u in ui]}


class _Factors:
    """Deduplicated factor table of one component."""

    def __init__(self):
        self.index = {}
        self.ints = []
        self.dbls = []
        self.special = {}       # factor index -> (variable, occurrence slot, cross-term flag) of a special term

    def add(self, var, kind, order=0, scale=1.0, scale2=0.0, mu=0.0, sigma=1.0, ident=None):
        """`ident` identifies a special term by its position (occurrence slot, cross-term flag) instead of by its
        centre/scale: those move with the data on reset(), and the tables (indices, blob sizes) must not."""
        key = (int(var), int(kind), int(order), float(scale), float(scale2)) + \
              ((float(mu), float(sigma)) if ident is None else ('st',) + tuple(ident))
        if key not in self.index:
            self.index[key] = len(self.ints)
            self.ints.append((int(var), int(kind), int(order), 0))
            self.dbls.append((float(scale), float(scale2), float(mu), float(sigma)))
            if ident is not None:
                self.special[len(self.ints) - 1] = (int(var), int(ident[0]), bool(ident[1]))
        return self.index[key]


class ComponentPlan:
    """Compiled tables of map component k (column c = k + skip_dimensions)."""

    def __init__(self, k, c, dtot, family, polyfunc, polyfunc_der, mon_spec, non_spec, special_terms,
                 linearization=None):
        self.k, self.c, self.dtot, self.family = k, c, dtot, family
        self.polyfunc, self.polyfunc_der = polyfunc, polyfunc_der
        cnt = np.zeros(dtot, dtype=int)
        self.mon_recs = [parse_entry(e, cnt, c, True, linearization) for e in mon_spec]
        cnt = np.zeros(dtot, dtype=int)
        self.non_recs = [parse_entry(e, cnt, c, False, linearization) for e in non_spec]
        # L-BFGS-B bounds of the separable fit: [0, inf) except constants (tm.py:1891-1892, 1925-1929)
        self.lb = np.array([-np.inf if r['type'] == 'const' else 0.0 for r in self.mon_recs])
        self.ub = np.full(len(self.mon_recs), np.inf)
        self.has_special = any(r['type'] == 'st' for r in self.mon_recs + self.non_recs)
        self.build(special_terms)

    def refresh_special(self, special_terms):
        """New centres / scales of the special terms into the existing double blob (the tables and the int blob do
        not depend on them).  Equivalent to build(special_terms), at a fraction of its cost."""
        for f, var, slot, cross, pos_fac, pos_ent in self._st_patch:
            d = special_terms[self.c]['cross-terms'] if cross else special_terms[self.c]
            mu, sg = float(d[var]['centers'][slot]), float(d[var]['scales'][slot])
            self.dblob[pos_fac + 2], self.dblob[pos_fac + 3] = mu, sg
            for pe in pos_ent:
                self.dblob[pe + 1], self.dblob[pe + 2] = mu, sg
        return self

    # ------------------------------------------------------------------ factors of one record
    def _st_params(self, rec, special_terms):
        d = special_terms[self.c]['cross-terms'] if rec['cross'] else special_terms[self.c]
        return float(d[rec['var']]['centers'][rec['slot']]), float(d[rec['var']]['scales'][rec['slot']])

    def _standard(self, rec, F, special_terms):
        """Factor list of a term in 'standard' mode (tm.py:885-1160)."""
        if rec['type'] == 'const':
            return [F.add(0, F_ONE)]
        if rec['type'] == 'st':
            mu, sg = self._st_params(rec, special_terms)
            return [F.add(rec['var'], _ST_KIND[rec['kind']], 0, 1.0, 0.0, mu, sg, ident=(rec['slot'], rec['cross']))]
        out = []
        for u, o in zip(rec['vars'], rec['orders']):
            if rec['hf']:
                out.append(F.add(u, F_POLY_HF, o, hf_normaliser(self.polyfunc, o)))
            else:
                out.append(F.add(u, F_POLY, o, 1.0))
        return out

    def _derivative(self, rec, F, special_terms, registry):
        """Factor list of d(term)/dx_c the way the reference writes it (tm.py:892-1258), including:
          * the first-wins dictionary of precalculated polynomials shared by 'HF' and plain factors
            of the same order on x_c (keys P_c_O_n / P_c_O_n_DER, tm.py:1169-1202, 1240);
          * the 'HF' derivative *assigning* the term string (tm.py:1245), which drops the factors
            that precede x_c.
        Returns (factor list, detected variable for the special-term grid or None)."""
        c = self.c
        if rec['type'] == 'const':
            return [F.add(0, F_ZERO)], None
        if rec['type'] == 'st':
            if rec['var'] != c:
                return [F.add(0, F_ZERO)], -1
            mu, sg = self._st_params(rec, special_terms)
            return [F.add(c, _ST_DKIND[rec['kind']], 0, 1.0, 0.0, mu, sg, ident=(rec['slot'], rec['cross']))], c
        if c not in rec['vars']:
            return [F.add(0, F_ZERO)], None
        out = []
        for u, o in zip(rec['vars'], rec['orders']):
            s = hf_normaliser(self.polyfunc, o) if rec['hf'] else 1.0
            if u != c:
                out.append(F.add(u, F_POLY_HF if rec['hf'] else F_POLY, o, s))
                continue
            registry.setdefault(('DER', u, o), s)
            if not rec['hf']:
                out.append(F.add(u, F_DPOLY, o, 1.0))
            else:
                registry.setdefault(('P', u, o), s)
                out = [F.add(u, F_DPOLY_HF, o, registry[('P', u, o)], registry[('DER', u, o)])]
        return out, None

    @staticmethod
    def _grid(recs, terms, dims, has_cross):
        """Special-term tensor grid of the monotone function (tm.py:1446-1483 / 2000-2037)."""
        if not has_cross:
            return terms
        st_idx = [i for i, r in enumerate(recs) if r['type'] == 'st']
        groups = {}
        for i in st_idx:
            groups.setdefault(dims[i], []).append(terms[i])
        keys = sorted(groups.keys())
        grid = [list(t) for t in groups[keys[0]]]
        for d in keys[1:]:
            grid = [a + b for a, b in itertools.product(grid, groups[d])]
        return [t for i, t in enumerate(terms) if i not in st_idx] + grid

    # ------------------------------------------------------------------ tables
    def build(self, special_terms):
        c = self.c
        F = _Factors()
        has_cross = 'cross-terms' in special_terms[c]
        non = [self._standard(r, F, special_terms) for r in self.non_recs]
        mon = self._grid(self.mon_recs, [self._standard(r, F, special_terms) for r in self.mon_recs],
                         [r['var'] if r['type'] == 'st' else None for r in self.mon_recs], has_cross)
        registry, dterms, ddims = {}, [], []
        for r in self.mon_recs:
            t, d = self._derivative(r, F, special_terms, registry)
            dterms.append(t)
            ddims.append(d)
        dmon = self._grid(self.mon_recs, dterms, ddims, has_cross)
        self.m_non, self.m_mon, self.m_dmon = len(non), len(mon), len(dmon)
        fi = np.asarray(F.ints, dtype=np.int32).reshape(-1, 4)
        fd = np.asarray(F.dbls, dtype=np.float64).reshape(-1, 4)

        # ---- nonmonotone terms: constants / univariate groups / multivariate rest
        const_idx, multi_idx, groups = [], [], {}
        for j, t in enumerate(non):
            if len(t) == 1 and fi[t[0], 1] == F_ONE:
                const_idx.append(j)
            elif len(t) == 1:
                groups.setdefault(int(fi[t[0], 0]), []).append((t[0], j))
            else:
                multi_idx.append(j)
        # dense groups: polynomial terms of one variable addressed by slot 2*order+hf (first occurrence of a
        # slot; duplicates and special terms go to the generic "slow" groups)
        dense_maxord = max([int(fi[f, 2]) for v in groups for f, _ in groups[v] if fi[f, 1] <= F_POLY_HF], default=0)
        stride = 2 * (dense_maxord + 1)
        dense_var, dense_idx, dense_scale, slow = [], [], [], {}
        for v in sorted(groups):
            idx_row, sc_row = -np.ones(stride, dtype=np.int32), np.zeros(stride)
            for f, j in groups[v]:
                slot = 2 * int(fi[f, 2]) + int(fi[f, 1] == F_POLY_HF)
                if fi[f, 1] <= F_POLY_HF and idx_row[slot] < 0:
                    idx_row[slot], sc_row[slot] = j, fd[f, 0]
                else:
                    slow.setdefault(v, []).append((f, j))
            used = np.nonzero(idx_row >= 0)[0]
            if len(used):
                dense_var.append((v, int(used.max() // 2), int(np.any(used % 2 == 1)), int(np.any(used % 2 == 0))))
                dense_idx.append(idx_row)
                dense_scale.append(sc_row)
        var_idx, var_ptr, ent_i, ent_d = [], [0], [], []
        for v in sorted(slow):
            ents = sorted(slow[v], key=lambda e: (0, fi[e[0], 2]) if fi[e[0], 1] <= F_POLY_HF else (1, 0))
            flags = 1 if any(fi[f, 1] == F_POLY_HF for f, _ in ents) else 0
            var_idx.append((v, flags))
            for f, j in ents:
                ent_i.append((int(fi[f, 1]), int(fi[f, 2]), j, 0))
                ent_d.append((fd[f, 0], fd[f, 2], fd[f, 3], 0.0))
            var_ptr.append(len(ent_i))
        touched = [int(fi[f, 0]) for t in non for f in t if fi[f, 1] not in (F_ONE, F_ZERO)]
        non_maxvar = 1 + max(touched) if touched else 0

        # ---- monotone terms: outer product over x_<c and one slot of x_c
        inner, outer = [], []
        for t in mon:
            inn = [f for f in t if fi[f, 0] == c and fi[f, 1] not in (F_ONE, F_ZERO)]
            out = [f for f in t if not (fi[f, 0] == c and fi[f, 1] not in (F_ONE, F_ZERO)) and fi[f, 1] != F_ONE]
            if len(inn) > 1:
                raise NotImplementedError('monotone term with several factors of x_%d' % c)
            inner.append(inn[0] if inn else None)
            outer.append(out)
        poly_inner = [f for f in inner if f is not None and fi[f, 1] <= F_POLY_HF]
        maxord = max([int(fi[f, 2]) for f in poly_inner], default=0)
        st_list = []
        for f in inner:
            if f is not None and fi[f, 1] > F_POLY_HF and f not in st_list:
                st_list.append(f)
        nst = len(st_list)
        nslot = 2 * (maxord + 1) + nst
        slot_terms = [[] for _ in range(nslot)]
        slot_scale = np.ones(nslot)
        for j, f in enumerate(inner):
            if f is None:
                s = 0
            elif fi[f, 1] == F_POLY:
                s = 2 * int(fi[f, 2])
                slot_scale[s] = fd[f, 0]
            elif fi[f, 1] == F_POLY_HF:
                s = 2 * int(fi[f, 2]) + 1
                slot_scale[s] = fd[f, 0]
            else:
                s = 2 * (maxord + 1) + st_list.index(f)
            slot_terms[s].append(j)
        has_plain = int(any(len(slot_terms[2 * o]) for o in range(maxord + 1)))
        has_hf = int(any(len(slot_terms[2 * o + 1]) for o in range(maxord + 1)))

        # ---- pack
        ib = [np.zeros(H_SIZE, dtype=np.int32)]
        db = []
        pos = {'i': H_SIZE, 'd': 0}

        def put_i(arr, align=1):
            arr = np.asarray(arr, dtype=np.int32).ravel()
            pad = (-pos['i']) % align
            if pad:
                ib.append(np.zeros(pad, dtype=np.int32))
                pos['i'] += pad
            off = pos['i']
            ib.append(arr)
            pos['i'] += arr.size
            return off

        def put_d(arr, align=1):
            arr = np.asarray(arr, dtype=np.float64).ravel()
            pad = (-pos['d']) % align
            if pad:
                db.append(np.zeros(pad))
                pos['d'] += pad
            off = pos['d']
            db.append(arr)
            pos['d'] += arr.size
            return off

        def csr(lists):
            ptr = np.cumsum([0] + [len(t) for t in lists])
            flat = [f for t in lists for f in t]
            return ptr, flat

        h = ib[0]
        h[H_MAGIC], h[H_DTOT], h[H_C], h[H_FAMILY], h[H_NFAC] = PLAN_MAGIC, self.dtot, c, self.family, len(fi)
        h[H_FAC_I] = put_i(fi, 4)
        h[H_D_FAC] = put_d(fd, 4)
        for (hm, hp, hf_), lists in (((H_M_NON, H_NON_PTR, H_NON_FAC), non), ((H_M_MON, H_MON_PTR, H_MON_FAC), mon),
                                     ((H_M_DMON, H_DMON_PTR, H_DMON_FAC), dmon)):
            ptr, flat = csr(lists)
            h[hm], h[hp], h[hf_] = len(lists), put_i(ptr), put_i(flat)
        h[H_NCONST], h[H_CONST_IDX] = len(const_idx), put_i(const_idx)
        h[H_NVARS], h[H_VAR_IDX], h[H_VAR_PTR] = len(var_idx), put_i(var_idx, 2), put_i(var_ptr)
        h[H_ENT_I] = put_i(ent_i, 4)
        h[H_D_ENT] = put_d(ent_d, 4)
        h[H_NMULTI], h[H_MULTI_IDX] = len(multi_idx), put_i(multi_idx)
        h[H_MAXORD], h[H_HAS_PLAIN], h[H_HAS_HF], h[H_NST], h[H_NSLOT] = maxord, has_plain, has_hf, nst, nslot
        ptr, flat = csr(slot_terms)
        h[H_SLOT_PTR], h[H_SLOT_TERM] = put_i(ptr), put_i(flat)
        ptr, flat = csr(outer)
        h[H_OUT_PTR], h[H_OUT_FAC] = put_i(ptr), put_i(flat)
        h[H_ST_FAC] = put_i(st_list)
        h[H_D_SLOT_SCALE] = put_d(slot_scale)
        h[H_D_REC] = put_d(np.zeros(1))
        h[H_NON_MAXVAR] = non_maxvar
        h[H_NDENSE], h[H_DENSE_MAXORD] = len(dense_var), dense_maxord
        h[H_NACTIVE] = sum(1 for t in slot_terms if len(t))
        h[H_NOUTFAC] = sum(len(t) for t in outer)
        h[H_DENSE_VAR] = put_i(dense_var, 4)
        h[H_DENSE_IDX] = put_i(np.concatenate(dense_idx) if dense_idx else [])
        h[H_D_DENSE_SCALE] = put_d(np.concatenate(dense_scale) if dense_scale else [])
        self.iblob = np.ascontiguousarray(np.concatenate(ib), dtype=np.int32)
        self.dblob = np.ascontiguousarray(np.concatenate(db), dtype=np.float64)
        self.maxord, self.nst, self.nslot = maxord, nst, nslot
        self.dense_maxord = dense_maxord
        # where the centres / scales of the special terms sit in the double blob: refresh_special() patches them in
        # place when the ensemble changes (reset(), tm.py:800) instead of recompiling the tables
        o_fac, o_ent = int(h[H_D_FAC]), int(h[H_D_ENT])
        ent_of = {}
        e = 0
        for v in sorted(slow):
            for f, _ in sorted(slow[v], key=lambda t: (0, fi[t[0], 2]) if fi[t[0], 1] <= F_POLY_HF else (1, 0)):
                ent_of.setdefault(int(f), []).append(e)
                e += 1
        self._st_patch = [(f, var, slot, cross, o_fac + 4 * f, [o_ent + 4 * q for q in ent_of.get(f, [])])
                          for f, (var, slot, cross) in F.special.items()]
        # host copies of the nonmonotone structure (the fused inverse packs its coefficient matrix from them)
        self.const_idx = list(const_idx)
        self.n_slow, self.n_multi = len(var_idx), len(multi_idx)
        self.dense_groups = [(int(dv[0]), np.asarray(di), np.asarray(ds))
                             for dv, di, ds in zip(dense_var, dense_idx, dense_scale)]
        self.non_maxvar = non_maxvar
        return self


def donor_map(nonmonotone):
    """{k: d}: d = the component with the longest nonmonotone term list that starts with component k's list (k itself
    if there is none, or if the list holds special terms -- those are placed per component, tm.py:2241-2330).  The
    Gram matrix of d's basis contains the one of k's as its leading block, so components that share a basis share one
    K-gram launch and one factorisation.  One pass, longest list first: a list's donor is always a root (a list that is
    no proper prefix of a longer one), so only roots are compared (list equality runs at C speed); the previous
    all-pairs search cost O(D^2 x terms) Python operations per fit (0.7 s at D = 256)."""
    D = len(nonmonotone)
    has_str = [any(type(e) == str for e in nonmonotone[k]) for k in range(D)]
    roots, dm = [], {}
    for k in sorted(range(D), key=lambda q: (-len(nonmonotone[q]), q)):
        spec = nonmonotone[k]
        dm[k] = k
        if has_str[k]:
            continue
        for r in roots:
            if nonmonotone[r][:len(spec)] == spec:
                dm[k] = r
                break
        else:
            roots.append(k)
    return dm


# ---------------------------------------------------------------------------------------------------------------------
# Packed operands of the fused kernels (host side, pure numpy: tested on the CPU against the oracle's basis)
# ---------------------------------------------------------------------------------------------------------------------
FUSED_CB = 16            # components per block of K-inv-fused (ttm_inverse_fused.cu: CB)
RECT_TC = 128            # components per tile of K-inv-rect (ttm_inverse_fused.cu: TC)


def fused_slots(plans):
    """Slot list (bit 2*order + hf of a dense group) shared by the components, or None when a component is outside the
    class of the fused kernels: nonmonotone terms = constants + per-variable groups of order <= 3."""
    if any(p.n_slow or p.n_multi or p.dense_maxord > 3 for p in plans):
        return None
    used = set()
    for p in plans:
        for _, idx_row, _ in p.dense_groups:
            used.update(int(q) for q in np.nonzero(idx_row >= 0)[0])
    if used - {2, 3, 4, 5, 6, 7}:
        return None
    return [2, 5, 7] if used <= {2, 5, 7} else [2, 3, 4, 5, 6, 7]


def fused_apack_doubles(ncomp, c0, ns):
    """Size of Apack (ttm_inverse_fused_apack_size): block b holds the rows v = 0 .. c0 + 16 b + 15."""
    nblk = (ncomp + FUSED_CB - 1) // FUSED_CB
    return (nblk * (c0 + FUSED_CB) + FUSED_CB * nblk * (nblk - 1) // 2) * FUSED_CB * ns


def rect_rpack_doubles(ncomp, rows, ns):
    """Size of Rpack (ttm_inverse_rect_rpack_size): [ceil(ncomp/128)][rows rounded up to 8][ns][128]."""
    return ((ncomp + RECT_TC - 1) // RECT_TC) * ((rows + 7) // 8 * 8) * ns * RECT_TC


def pack_fused_operands(plans, first_col, rect_rows, want_apack=True):
    """Scatter indices of coefficient*scale into the packed operands of K-inv-fused (Apack) and K-inv-rect / K-map-rect
    (Rpack) for the components `plans` (component j solves / evaluates column first_col + j).

    rect_rows: number of leading variables whose contribution goes into Rpack (0: none; the conditioning width E for
    the split inverse; Dtot - 1 for the forward map).  Returns None when the class does not fit, else a dict with
      ns, slots, src (index into the concatenated coefficients_nonmon of the components), sc (scale),
      dst (index into Apack, if want_apack), rdst / rkeep (index into Rpack for the entries with v < rect_rows),
      const_src / const_ptr (CSR of the constant terms of each component: a0_j = sum of those coefficients)."""
    slots = fused_slots(plans)
    if slots is None or any(p.c != first_col + j for j, p in enumerate(plans)):
        return None
    ns, CB = len(slots), FUSED_CB
    rp = (rect_rows + 7) // 8 * 8
    dst, src, sc, rdst, const_src, const_ptr, off = [], [], [], [], [], [0], 0
    for j, p in enumerate(plans):
        b, jj = divmod(j, CB)
        row0 = b * (first_col + CB) + CB * b * (b - 1) // 2
        for v, idx_row, sc_row in p.dense_groups:
            if v >= first_col + j:
                return None                                                # not a triangular dependency
            for q, sl in enumerate(slots):
                if sl < len(idx_row) and idx_row[sl] >= 0:
                    dst.append(((row0 + v) * CB + jj) * ns + q)
                    src.append(off + int(idx_row[sl]))
                    sc.append(float(sc_row[sl]))
                    rdst.append((((j // RECT_TC) * rp + v) * ns + q) * RECT_TC + j % RECT_TC if v < rect_rows else -1)
        const_src += [off + int(q) for q in p.const_idx]
        const_ptr.append(len(const_src))
        off += p.m_non
    rdst = np.asarray(rdst, dtype=np.int64)
    rkeep = rdst >= 0
    out = {'ns': ns, 'slots': slots, 'src': np.asarray(src, dtype=np.int64), 'sc': np.asarray(sc, dtype=np.float64),
           'rdst': rdst[rkeep], 'rkeep': rkeep, 'const_src': np.asarray(const_src, dtype=np.int64),
           'const_ptr': np.asarray(const_ptr, dtype=np.int64)}
    if want_apack:
        out['dst'] = np.asarray(dst, dtype=np.int64)
    return out
