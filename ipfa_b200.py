"""Import alias: ``import ipfa_b200`` == the package directory
``iterative-pseudo-forced-alignment-ctc_b200`` (whose name is not a Python identifier)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200")
sys.modules[__name__] = _pkg
