"""Import alias for the ``raytracer-3_b200/`` package directory."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raytracer-3_b200")
_spec = importlib.util.spec_from_file_location("rt3_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rt3_b200"] = _mod
_spec.loader.exec_module(_mod)
