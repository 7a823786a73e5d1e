"""Import alias: the package directory is named `vplines-slam_b200` (not a valid
Python identifier); `import vplines_slam_b200` resolves to it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_mod = importlib.import_module("vplines-slam_b200")
sys.modules[__name__] = _mod
