"""Importable alias of the ``lighting-asr_b200`` package (whose directory name is not a valid
Python identifier).  ``import lasr_b200`` / ``"lasr_b200:GpuFbankFrontend"`` in a LASR
config.yaml resolve to the same module objects."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("lighting-asr_b200")
sys.modules[__name__] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("lighting-asr_b200."):
        sys.modules["lasr_b200." + _name.split(".", 1)[1]] = _mod
