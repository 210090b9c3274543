"""Importable alias of the product package (its directory name, fixed by the project layout, has a hyphen):
``import b200rl`` == ``importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
