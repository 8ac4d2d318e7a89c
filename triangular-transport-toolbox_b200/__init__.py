"""B200-native hot path of the Triangular Transport Toolbox.

The directory name carries a hyphen, so it is loaded by path: the root-level `transport_map.py`
(the drop-in module: `from transport_map import *`) registers it as `ttt_b200`.
"""
from .transport_map import transport_map  # noqa: F401

__all__ = ['transport_map']
