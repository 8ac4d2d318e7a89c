"""The constants of the kernels' exp (csrc/ttm_exp.cuh, ttm_exp_tab64.h / ttm_exp_tab32.h), checked on the CPU: the
tables must hold the correctly rounded 2^(j/E) * 2^-1021, and the reduction + polynomial restated in numpy must stay
within the error the header states.  (The kernels use FMAs; numpy's separate roundings add ~1e-16.)"""
import os
import re

import numpy as np
import pytest

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'triangular-transport-toolbox_b200', 'csrc')


def table(entries):
    text = open(os.path.join(CSRC, 'ttm_exp_tab%d.h' % entries)).read()
    lo = [int(w, 16) for w in re.findall(r'0x([0-9a-f]{8})u', text.split('_lo[%d]' % entries)[1].split('};')[0])]
    hi = [int(w, 16) for w in re.findall(r'0x([0-9a-f]{8})u', text.split('_hi[%d]' % entries)[1].split('};')[0])]
    assert len(lo) == entries and len(hi) == entries
    bits = (np.asarray(hi, dtype=np.uint64) << np.uint64(32)) | np.asarray(lo, dtype=np.uint64)
    return bits.view(np.float64)


def macros(variant):
    text = open(os.path.join(CSRC, 'ttm_exp.cuh')).read()
    block = text.split('#if TTM_EXP_VARIANT == 64')[2]           # [0]: header include switch, [2]: the constants
    part = block.split('#else')[0] if variant == 64 else block.split('#else')[1].split('#endif')[0]
    out = {}
    for name, val in re.findall(r'#define TTM_E32_(\w+) ([-0-9.e]+)\s', part):
        out[name] = float(val)
    return out


@pytest.mark.parametrize('entries', [64, 32])
def test_table_holds_correctly_rounded_powers(entries):
    mp = pytest.importorskip('mpmath')
    mp.mp.dps = 60
    tab = table(entries)
    for j in range(entries):
        exact = mp.mpf(2) ** (mp.mpf(j) / entries)
        got = mp.mpf(float(tab[j])) * mp.mpf(2) ** 1021
        assert abs(got - exact) <= abs(exact) * mp.mpf(2) ** -53, j     # half an ulp


@pytest.mark.parametrize('variant,bound', [(64, 7e-15), (32, 1e-15)])
def test_reduction_and_polynomial_stay_within_the_stated_error(variant, bound):
    m = macros(variant)
    E = variant
    assert abs(m['K'] - E / np.log(2)) <= 1e-13 and abs(m['C'] - np.log(2) / E) <= 1e-17
    tab = table(E) * 2.0 ** 1021
    x = np.concatenate((np.linspace(-40.0, 40.0, 400001), np.linspace(-1e-3, 1e-3, 20001)))
    magic = 6755399441055744.0
    t = x * m['K'] + magic
    nf = t - magic
    k = nf.astype(np.int64)
    r = x - nf * m['C']
    assert np.max(np.abs(r)) <= np.log(2) / (2 * E) * (1 + 1e-9)
    if variant == 64:
        q = (m['Q2'] * r + m['Q1']) * r + m['Q0']
    else:
        q = ((m['Q3'] * r + m['Q2']) * r + m['Q1']) * r + m['Q0']
    p = ((q * r + 1.0) * r + 1.0) * tab[k % E]
    got = np.ldexp(p, (k // E).astype(np.int64))
    ref = np.exp(x.astype(np.longdouble))
    err = float(np.max(np.abs((got.astype(np.longdouble) - ref) / ref)))
    # the single-constant reduction perturbs the argument by up to |x| 2^-53 (rounding of ln2/E), on top of the polynomial
    assert err <= bound + 40.0 * 2.0 ** -53, err
