"""
Coefficient persistence in the format the reference's examples use, plus a chronicle-style fit log.

The examples pickle `{'coeffs_mon': tm.coeffs_mon, 'coeffs_nonmon': tm.coeffs_nonmon}` after `optimize()` and write
the lists back into a fresh map object on the next run (example_01.py:215-231); the adaptation routines pickle a
`chronicle` dictionary keyed by component and iteration (tm.py:4647-4660, :4703-4711, :4943-4950).  These helpers
read and write exactly those layouts, so files move freely between the reference and this package.
"""

import copy
import pickle

import numpy as np


def coefficient_dict(tm):
    """The examples' dictionary: lists of float64 vectors, one per map component."""
    return {'coeffs_mon': [np.array(c, dtype=np.float64) for c in tm.coeffs_mon],
            'coeffs_nonmon': [np.array(c, dtype=np.float64) for c in tm.coeffs_nonmon]}


def save_coefficients(tm, path):
    """pickle.dump of the examples' dictionary (example_01.py:215-224)."""
    with open(path, 'wb') as f:
        pickle.dump(coefficient_dict(tm), f)


def load_coefficients(tm, path):
    """Write pickled coefficients into the map without optimising (example_01.py:226-231).  Accepts files written by
    the reference's examples or by save_coefficients; checks the term counts against the map."""
    with open(path, 'rb') as f:
        d = pickle.load(f)
    cm, cn = d['coeffs_mon'], d['coeffs_nonmon']
    if len(cm) != tm.D or len(cn) != tm.D:
        raise ValueError('coefficient file holds %d/%d components, the map has %d' % (len(cm), len(cn), tm.D))
    for k in range(tm.D):
        if len(cm[k]) != len(tm.coeffs_mon[k]) or len(cn[k]) != len(tm.coeffs_nonmon[k]):
            raise ValueError('component %d: file has %d monotone / %d nonmonotone coefficients, the map %d / %d'
                             % (k, len(cm[k]), len(cn[k]), len(tm.coeffs_mon[k]), len(tm.coeffs_nonmon[k])))
    tm.coeffs_mon = [np.array(c, dtype=np.float64) for c in cm]
    tm.coeffs_nonmon = [np.array(c, dtype=np.float64) for c in cn]
    return tm


class Chronicle(dict):
    """Fit log in the layout of the reference's adaptation chronicle: chronicle[k][iteration] = record with the
    component's term lists and coefficients (tm.py:4703-4711), extended with the optimizer's counters."""

    def record(self, tm, k, iteration=None, **extra):
        per_k = self.setdefault(int(k), {})
        it = len(per_k) if iteration is None else int(iteration)
        rec = {'monotone': copy.deepcopy(tm.monotone[k]), 'nonmonotone': copy.deepcopy(tm.nonmonotone[k]),
               'coeffs_nonmon': np.array(tm.coeffs_nonmon[k], dtype=np.float64),
               'coeffs_mon': np.array(tm.coeffs_mon[k], dtype=np.float64)}
        rec.update(extra)
        per_k[it] = rec
        return rec

    def save(self, path='dictionary_adaptation_chronicle.p'):
        """Same default file name as the reference (tm.py:4948-4950)."""
        with open(path, 'wb') as f:
            pickle.dump(dict(self), f)

    @staticmethod
    def load(path='dictionary_adaptation_chronicle.p'):
        with open(path, 'rb') as f:
            c = Chronicle()
            c.update(pickle.load(f))
            return c
