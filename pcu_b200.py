"""Importable alias for the package directory `point-cloud-utils_b200/` (whose name, fixed by the
repository layout, is not a valid Python identifier).

    import pcu_b200 as pcu
    pcu.chamfer_distance(x, y)
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("point-cloud-utils_b200")
sys.modules[__name__] = _pkg
