"""Importable alias: the package directory carries the (hyphenated) repository
name, which `import` cannot spell.  `import sbce` gives the same module."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
PACKAGE = "semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200"
_pkg = importlib.import_module(PACKAGE)
sys.modules[__name__] = _pkg
