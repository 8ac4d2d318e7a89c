"""
Map adaptation on top of the CUDA path.

The reference's two structure searches (`adapt_map`, tm.py:373-657, and `adaptation_cross_terms`, :4575-4950) only
need four things from the hot path: re-compilation of the term tables of the whole map or of one component
(`function_constructor_alternative(k)`, :1263), `optimize()`, `map()` and `objective_function()`.  Here those run on
the GPU (term tables are recompiled on the host in microseconds, no source generation), while the search logic --
normality tests, precision/correlation thresholds, the multi-index frontier -- stays host Python, restated from
the reference's algorithm:

separable search (:405-640)
    1. marginal stage: start from S_k = a_0 + b_0 x_k; fit; every component whose output fails the Shapiro-Wilk
       test (p < threshold_sw) gets one more integrated-RBF term 'iRBF k'; repeat until all pass or the order cap.
    2. off-diagonal stage: fit; for every pair (k, j < k) whose |precision| (first pass) or correlation (later
       passes) of the map output exceeds threshold_prec, raise the order of x_j in component k by one
       (plain linear term first, Hermite functions afterwards); pairs below the threshold are frozen.
cross-term search (:4575-4950), integrated-rectifier maps
    per component: a multi-index set grown greedily; candidates = reduced margin of the current set; each candidate
    is scored by a one-sided finite difference of the objective when its coefficient is switched on; the best one
    joins the set and the component is re-fitted (L-BFGS-B on the objective, like the reference).

Reference quirks kept because they define the result: the marginal stage indexes the monotone variable by k (not
k + skip_dimensions) but counts orders in column k + skip_dimensions; the off-diagonal loop never revisits a frozen
pair; the cross-term re-fit uses scipy's finite-difference gradient (the reference passes no `jac`).
"""

import copy

import numpy as np


def _refit(tm, monotone, nonmonotone):
    tm.monotone, tm.nonmonotone = copy.deepcopy(monotone), copy.deepcopy(nonmonotone)
    tm.function_constructor_alternative()
    tm.precalculate()
    tm.optimize()
    return tm.map()


def _standardised_abs(M):
    M = np.abs(M)
    d = np.sqrt(np.diag(M))
    return M / d[np.newaxis, :] / d[:, np.newaxis]


def adapt_separable(tm, maxorder_mon=10, maxorder_nonmon=10, threshold_sw=0.1, threshold_prec=0.1, map_finished=None):
    """tm.py:405-640."""
    import scipy.stats
    D, skip = tm.D, tm.skip_dimensions
    nonmonotone = [[[]] for _ in range(D)]
    monotone = [[[k]] for k in np.arange(D)]
    maporders = np.zeros((D, D), dtype=int)
    np.fill_diagonal(maporders, 1)
    pvals = np.zeros((maxorder_mon, D))
    gaussian = np.zeros(D, dtype=bool)
    it = 0
    while True:                                              # ---- marginal stage
        it += 1
        Z = _refit(tm, monotone, nonmonotone)
        p = np.array([scipy.stats.shapiro(Z[:, k]).pvalue for k in range(D)])
        pvals[it - 1, :] = p
        gaussian[p >= threshold_sw] = True
        for k in np.where(~gaussian)[0]:
            if maporders[k, k + skip] < maxorder_mon:
                maporders[k, k + skip] += 1
                monotone[k] += ['iRBF ' + str(k)]
        if gaussian.all() or it >= maxorder_mon - 1:
            break
    tm.pvals_mat = pvals
    tm.covmat = _standardised_abs(np.cov(Z.T))
    tm.precmat = _standardised_abs(np.linalg.inv(np.cov(Z.T)))
    if map_finished is None:
        map_finished = np.zeros((D, D), dtype=bool)
    precmat_list = [tm.precmat.copy()]
    it = 0
    while True:                                              # ---- off-diagonal stage
        it += 1
        Z = _refit(tm, monotone, nonmonotone)
        stop = False
        try:
            prec = _standardised_abs(np.linalg.inv(np.cov(Z.T))) if it == 1 else np.corrcoef(Z.T)
            for k in range(D):
                for j in range(k):
                    if prec[k, j] > threshold_prec and not map_finished[k, j]:
                        maporders[k, j] += 1
                        o = int(maporders[k, j])
                        nonmonotone[k].append([j] * o if o == 1 else [j] * o + ['HF'])
                    else:
                        map_finished[k, j] = True
                nonmonotone[k].sort()
            precmat_list.append(prec.copy())
        except Exception:                                    # the reference stops on any failure here (:612-616)
            stop = True
        if stop or map_finished.sum() >= D * (D - 1) / 2:
            break
        if it >= maxorder_nonmon:
            print("WARNING: Map adaptation stopped at maximum number of iterations.")
            break
    tm.precmat_list = precmat_list
    _refit(tm, monotone, nonmonotone)
    tm.maporders = maporders
    return tm


def _cell_term(tm, cell):
    term = []
    for v, order in enumerate(cell):
        term += [int(v)] * int(order)
    if tm.polynomial_type.lower() == 'hermite function' and term:
        term += ['HF']
    return term


def _component_from_cells(tm, mim):
    """Term lists of one component from its multi-index matrix: cells with a positive order in the last variable
    are monotone terms, the others nonmonotone; negative entries mark proposed cells (tm.py:4594-4640)."""
    cells = np.asarray(np.where(mim != 0)).T
    monotone, nonmonotone, proposed, original = [], [], [], []
    for n, cell in enumerate(cells):
        (proposed if mim[tuple(cell)] < 0 else original).append(n)
        (monotone if cell[-1] > 0 else nonmonotone).append(_cell_term(tm, cell))
    return monotone, nonmonotone, proposed, original


def adapt_cross_terms(tm, increment=1e-6, chronicle=False):
    """tm.py:4575-4950."""
    from scipy.optimize import minimize
    from .persistence import Chronicle
    log = Chronicle() if chronicle else None
    nmax = tm.adaptation_max_order + 1
    for k in range(tm.D):
        nv = k + 1 + tm.skip_dimensions
        mim = np.zeros((nmax,) * nv, dtype=int)
        mim[(0,) * nv] = 1                                   # constant
        mim[(0,) * (nv - 1) + (1,)] = 1                      # linear in x_k
        tm.multi_index_matrix = mim
        div = len(tm.coeffs_nonmon[k])
        x0 = np.concatenate((tm.coeffs_nonmon[k], tm.coeffs_mon[k]))
        opt = minimize(method='BFGS', fun=tm.objective_function, jac=tm.objective_function_jacobian, x0=x0, args=(k, div))
        coeffs = opt.x.copy()
        tm.coeffs_nonmon[k], tm.coeffs_mon[k] = coeffs[:div].copy(), coeffs[div:].copy()
        if log is not None:
            log.record(tm, k, 0, multi_index_matrix=mim.copy())
        it = 0
        while True:
            it += 1
            # reduced margin: neighbours of the active cells, counted once per active neighbour ...
            for cell in np.asarray(np.where(mim > 0)).T:
                for ax in range(nv):
                    for step in (-1, 1):
                        idx = list(cell)
                        idx[ax] += step
                        if 0 <= idx[ax] < nmax and mim[tuple(idx)] <= 0:
                            mim[tuple(idx)] -= 1
            prop = np.asarray(np.where(mim < 0)).T
            if len(prop) == 0:
                break
            for cell in prop:                                # ... plus once per zero index (boundary faces)
                mim[tuple(cell)] -= int(np.sum(cell == 0))
            prop = np.asarray(np.where(mim <= -nv)).T        # admissible: all backward neighbours present
            coeffs = np.concatenate((tm.coeffs_nonmon[k], tm.coeffs_mon[k]))
            obj_ref = tm.objective_function(coeffs=coeffs, k=k, div=div)
            grads = np.zeros(len(prop))
            for n, cell in enumerate(prop):
                mim[mim < 0] = 0
                mim[tuple(cell)] = -1
                mon, non, _, orig = _component_from_cells(tm, mim)
                tm.monotone[k], tm.nonmonotone[k] = copy.deepcopy(mon), copy.deepcopy(non)
                c_new = np.ones(len(non) + len(mon)) * tm.coeffs_init + increment
                c_new[orig] = coeffs
                div = len(non)
                tm.function_constructor_alternative(k=k)
                grads[n] = (tm.objective_function(coeffs=c_new, k=k, div=div) - obj_ref) / increment
            best = np.where(np.abs(grads) == np.max(np.abs(grads)))[0][0]
            mim[mim < 0] = 0
            added = prop[best]
            mim[tuple(added)] = -1
            mon, non, _, orig = _component_from_cells(tm, mim)
            mim[tuple(added)] = 1
            c_new = np.ones(len(non) + len(mon)) * tm.coeffs_init
            c_new[orig] = coeffs
            div = len(non)
            tm.monotone[k], tm.nonmonotone[k] = copy.deepcopy(mon), copy.deepcopy(non)
            tm.function_constructor_alternative(k=k)
            opt = minimize(method='L-BFGS-B', fun=tm.objective_function, x0=c_new, args=(k, div))
            coeffs = opt.x.copy()
            tm.coeffs_nonmon[k], tm.coeffs_mon[k] = coeffs[:div].copy(), coeffs[div:].copy()
            if log is not None:
                log.record(tm, k, it, multi_index_matrix=mim.copy())
            if it >= tm.adaptation_max_iterations:
                break
    if log is not None:
        log.save('dictionary_adaptation_chronicle.p')
    return tm
