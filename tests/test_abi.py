"""C-ABI library: loads, and exports every entry point include/ttm.h declares (no compute calls: CPU box)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'triangular-transport-toolbox_b200', 'libttm.so')


def declared():
    text = open(os.path.join(ROOT, 'include', 'ttm.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ttm_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_hot_path_entry_points():
    names = declared()
    for must in ('ttm_objgrad_ir', 'ttm_basis_eval', 'ttm_gram', 'ttm_sep_objgrad', 'ttm_inverse_table',
                 'ttm_inverse_bisect', 'ttm_eval_s_ir', 'ttm_sep_eval', 'ttm_standardize_transpose'):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared() if not hasattr(lib, s)]
    assert not missing, missing
    lib.ttm_version.restype = ctypes.c_int
    assert lib.ttm_version() >= 100


def test_binding_signatures_cover_the_header():
    import transport_map  # noqa: F401
    from ttt_b200 import binding
    L = binding.lib()
    for s in declared():
        assert hasattr(L, s)


def test_product_fails_loudly_without_cuda():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from transport_map import transport_map
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        transport_map(X=np.zeros((8, 1)), monotone=[[[0]]], nonmonotone=[[[]]])
