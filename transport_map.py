"""
Drop-in module: `from transport_map import *` keeps working (reference README.md:5,
example_01.py:12) and now yields the CUDA-backed class from `triangular-transport-toolbox_b200/`.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'triangular-transport-toolbox_b200')
_NAME = 'ttt_b200'

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, '__init__.py'),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

from ttt_b200.transport_map import transport_map  # noqa: E402,F401

__all__ = ['transport_map']
