"""Importable alias for the package directory `ai-interior-image-classifier_b200/` (hyphens are not valid in an
`import` statement):  `import iic_b200`  ==  importlib.import_module("ai-interior-image-classifier_b200")."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("ai-interior-image-classifier_b200")
sys.modules[__name__] = _pkg
