"""
Golden fixtures of the two structure searches, produced by the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_adapt.py

  adapt_separable.npz    adapt_map (tm.py:373-657, separable branch) on a seeded 3-D ensemble
  adapt_cross_terms.npz  adaptation_cross_terms (tm.py:4575-4950) on a seeded 2-D ensemble

Stored: the adapted term lists (repr), maporders / multi-index matrix, fitted coefficients, map() of the ensemble.
"""

import copy
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, '/root/reference')

import scipy                                                     # noqa: E402
from transport_map import transport_map                          # noqa: E402  (the reference)
from cases import adapt_separable_case, adapt_cross_case         # noqa: E402

VERS = np.asarray('numpy %s scipy %s' % (np.__version__, scipy.__version__))


def dump(name, tm, extra):
    out = {'monotone': np.asarray(repr(tm.monotone)), 'nonmonotone': np.asarray(repr(tm.nonmonotone)), '_versions': VERS}
    for k in range(tm.D):
        out['coeffs_mon_%d' % k] = np.array(tm.coeffs_mon[k])
        out['coeffs_nonmon_%d' % k] = np.array(tm.coeffs_nonmon[k])
    out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, repr(tm.monotone), repr(tm.nonmonotone))


def main():
    X, kw, call = adapt_separable_case()
    tm = transport_map(X=copy.copy(X), **kw)
    with redirect_stdout(io.StringIO()):
        tm.adapt_map(**call)
    dump('adapt_separable', tm, {'maporders': tm.maporders, 'map_train': tm.map()})

    X, kw, call = adapt_cross_case()
    tm = transport_map(X=copy.copy(X), **kw)
    with redirect_stdout(io.StringIO()):
        tm.adaptation_cross_terms(**call)
    # the cross-term search leaves Psi of the last component in place; map() of the training set uses the functions
    dump('adapt_cross_terms', tm, {'multi_index_matrix': tm.multi_index_matrix, 'map_train': tm.map(copy.copy(X))})


if __name__ == '__main__':
    main()
