"""numpy interpreter of the compiled term tables (plan blobs) -- test infrastructure.

Lets the CPU suite check the host-side term compiler (`plan.py`) against the oracle without a GPU:
the CUDA kernels read exactly these tables (layout: csrc/ttm_common.cuh)."""

import numpy as np
import scipy.special

import transport_map as _shim                          # noqa: F401  registers the package as ttt_b200
from ttt_b200 import plan as PL

_FAMS = {PL.FAM_POWER: np.polynomial.polynomial.Polynomial, PL.FAM_HERMITE: np.polynomial.hermite.Hermite,
         PL.FAM_HERMITE_E: np.polynomial.hermite_e.HermiteE, PL.FAM_CHEBYSHEV: np.polynomial.chebyshev.Chebyshev,
         PL.FAM_LAGUERRE: np.polynomial.laguerre.Laguerre, PL.FAM_LEGENDRE: np.polynomial.legendre.Legendre}


def factor(kind, order, scale, scale2, mu, sg, family, x):
    P = _FAMS[family]
    e = [0.] * order + [1.]
    s2 = np.sqrt(2)
    if kind == PL.F_POLY:
        return scale * P(e)(x)
    if kind == PL.F_POLY_HF:
        return scale * P(e)(x) * np.exp(-x ** 2 / 4)
    if kind == PL.F_DPOLY:
        return scale * P(e).deriv()(x)
    if kind == PL.F_DPOLY_HF:
        return -0.5 * np.exp(-x ** 2 / 4) * (x * scale * P(e)(x) - 2 * scale2 * P(e).deriv()(x))
    u = (x - mu) / (s2 * sg)
    if kind == PL.F_RBF:
        return np.exp(-((x - mu) / sg) ** 2 / 2) / (sg * np.sqrt(2 * np.pi))
    if kind == PL.F_IRBF:
        return (1 + scipy.special.erf(u)) / 2
    if kind == PL.F_LET:
        return ((x - mu) * (1 - scipy.special.erf(u)) - sg * np.sqrt(2 / np.pi) * np.exp(-u ** 2)) / 2
    if kind == PL.F_RET:
        return ((x - mu) * (1 + scipy.special.erf(u)) + sg * np.sqrt(2 / np.pi) * np.exp(-u ** 2)) / 2
    if kind == PL.F_DRBF:
        return -(x - mu) / (np.sqrt(2 * np.pi) * sg ** 3) * np.exp(-((x - mu) / sg) ** 2 / 2)
    if kind == PL.F_DIRBF:
        return np.exp(-(x - mu) ** 2 / (2 * sg ** 2)) / (np.sqrt(2 * np.pi) * sg)
    if kind == PL.F_DLET:
        return (1 - scipy.special.erf(u)) / 2
    if kind == PL.F_DRET:
        return (1 + scipy.special.erf(u)) / 2
    if kind == PL.F_ONE:
        return np.ones_like(x)
    return np.zeros_like(x)


class PlanInterp:
    def __init__(self, iblob, dblob):
        self.ib, self.db = iblob, dblob
        h = iblob
        assert h[PL.H_MAGIC] == PL.PLAN_MAGIC
        self.h = h
        nf = h[PL.H_NFAC]
        self.fi = iblob[h[PL.H_FAC_I]:h[PL.H_FAC_I] + 4 * nf].reshape(nf, 4)
        self.fd = dblob[h[PL.H_D_FAC]:h[PL.H_D_FAC] + 4 * nf].reshape(nf, 4)
        self.family = int(h[PL.H_FAMILY])

    def fac(self, f, X):
        v, kind, order, _ = (int(t) for t in self.fi[f])
        sc, sc2, mu, sg = self.fd[f]
        return factor(kind, order, sc, sc2, mu, sg, self.family, X[:, v])

    def csr(self, hm, hp, hf):
        m = int(self.h[hm])
        ptr = self.ib[self.h[hp]:self.h[hp] + m + 1]
        return m, ptr, self.ib[self.h[hf]:]

    def terms(self, which, X):
        hm, hp, hf = {0: (PL.H_M_NON, PL.H_NON_PTR, PL.H_NON_FAC), 1: (PL.H_M_MON, PL.H_MON_PTR, PL.H_MON_FAC),
                      2: (PL.H_M_DMON, PL.H_DMON_PTR, PL.H_DMON_FAC)}[which]
        m, ptr, flat = self.csr(hm, hp, hf)
        cols = []
        for j in range(m):
            v = None
            for q in range(ptr[j], ptr[j + 1]):
                f = self.fac(int(flat[q]), X)
                v = f if v is None else v * f
            cols.append(v)
        return np.stack(cols, axis=-1) if cols else None

    def nonmon_structured(self, X):
        """Psi_non rebuilt from the constant list, the per-variable entry groups and the multivariate rest
        (what the fused kernel's sweep reads)."""
        h = self.h
        m = int(h[PL.H_M_NON])
        out = np.full((X.shape[0], m), np.nan)
        for j in self.ib[h[PL.H_CONST_IDX]:h[PL.H_CONST_IDX] + h[PL.H_NCONST]]:
            out[:, j] = 1.0
        nd, dmax = int(h[PL.H_NDENSE]), int(h[PL.H_DENSE_MAXORD])
        stride = 2 * (dmax + 1)
        dvar = self.ib[h[PL.H_DENSE_VAR]:h[PL.H_DENSE_VAR] + 4 * nd].reshape(nd, 4)
        didx = self.ib[h[PL.H_DENSE_IDX]:h[PL.H_DENSE_IDX] + nd * stride].reshape(nd, stride)
        dsc = self.db[h[PL.H_D_DENSE_SCALE]:h[PL.H_D_DENSE_SCALE] + nd * stride].reshape(nd, stride)
        P = _FAMS[self.family]
        for g in range(nd):
            col, gmax, has_hf, has_plain = (int(t) for t in dvar[g])
            x = X[:, col]
            assert gmax <= dmax and np.all(didx[g, 2 * (gmax + 1):] < 0) and np.all(didx[g, :2] < 0)
            for s in range(stride):
                j = int(didx[g, s])
                if j < 0:
                    continue
                o, hf = divmod(s, 2)
                assert (has_hf if hf else has_plain)
                out[:, j] = dsc[g, s] * P([0.] * o + [1.])(x) * (np.exp(-x ** 2 / 4) if hf else 1.0)
        nv = int(h[PL.H_NVARS])
        var = self.ib[h[PL.H_VAR_IDX]:h[PL.H_VAR_IDX] + 2 * nv].reshape(nv, 2)
        ptr = self.ib[h[PL.H_VAR_PTR]:h[PL.H_VAR_PTR] + nv + 1]
        ne = int(ptr[-1]) if nv else 0
        ei = self.ib[h[PL.H_ENT_I]:h[PL.H_ENT_I] + 4 * ne].reshape(ne, 4)
        ed = self.db[h[PL.H_D_ENT]:h[PL.H_D_ENT] + 4 * ne].reshape(ne, 4)
        for g in range(nv):
            for e in range(ptr[g], ptr[g + 1]):
                kind, order, j, _ = (int(t) for t in ei[e])
                out[:, j] = factor(kind, order, ed[e, 0], 0.0, ed[e, 1], ed[e, 2], self.family, X[:, var[g, 0]])
        full = self.terms(0, X)
        for j in self.ib[h[PL.H_MULTI_IDX]:h[PL.H_MULTI_IDX] + h[PL.H_NMULTI]]:
            out[:, j] = full[:, j]
        return out

    def mon_structured(self, X):
        """Psi_mon rebuilt as outer product x slot basis (what the node loop uses)."""
        h = self.h
        m, maxord, nst = int(h[PL.H_M_MON]), int(h[PL.H_MAXORD]), int(h[PL.H_NST])
        nslot = int(h[PL.H_NSLOT])
        assert nslot == 2 * (maxord + 1) + nst
        sptr = self.ib[h[PL.H_SLOT_PTR]:h[PL.H_SLOT_PTR] + nslot + 1]
        sterm = self.ib[h[PL.H_SLOT_TERM]:]
        optr = self.ib[h[PL.H_OUT_PTR]:h[PL.H_OUT_PTR] + m + 1]
        ofac = self.ib[h[PL.H_OUT_FAC]:]
        stf = self.ib[h[PL.H_ST_FAC]:h[PL.H_ST_FAC] + nst]
        scale = self.db[h[PL.H_D_SLOT_SCALE]:h[PL.H_D_SLOT_SCALE] + nslot]
        xc = X[:, int(h[PL.H_C])]
        out = np.full((X.shape[0], m), np.nan)
        P = _FAMS[self.family]
        for s in range(nslot):
            if s < 2 * (maxord + 1):
                o, hf = divmod(s, 2)
                base = P([0.] * o + [1.])(xc) * (np.exp(-xc ** 2 / 4) if hf else 1.0)
                if sptr[s + 1] > sptr[s]:
                    assert (int(h[PL.H_HAS_HF]) if hf else int(h[PL.H_HAS_PLAIN]))
            else:
                f = int(stf[s - 2 * (maxord + 1)])
                base = self.fac(f, X)
            for jj in range(sptr[s], sptr[s + 1]):
                j = int(sterm[jj])
                u = np.ones(X.shape[0])
                for q in range(optr[j], optr[j + 1]):
                    u = u * self.fac(int(ofac[q]), X)
                out[:, j] = u * scale[s] * base
        return out
