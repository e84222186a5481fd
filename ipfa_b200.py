"""Import alias: ``import ipfa_b200`` == the package directory
``iterative-pseudo-forced-alignment-ctc_b200`` (whose name is not a Python identifier).
Submodules are registered under both names so ``from ipfa_b200.ctc_segmentation import …``
yields the same module objects as the hyphenated import."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_PKG = "iterative-pseudo-forced-alignment-ctc_b200"
_pkg = importlib.import_module(_PKG)
for _name in ("_lib", "build", "ops", "ctc_segmentation", "hostglue", "anchor", "words", "sharding", "stub_asr", "sweep"):
    _mod = importlib.import_module(f"{_PKG}.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    setattr(_pkg, _name, _mod)
sys.modules[__name__] = _pkg
