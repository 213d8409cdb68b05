"""Importable alias for the package directory `arc-alkali-rydberg-calculator_b200/`
(a hyphenated directory name cannot be imported directly).  `import arc_b200` loads
that directory as the package `arc_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "arc-alkali-rydberg-calculator_b200")
_spec = importlib.util.spec_from_file_location(
    "arc_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["arc_b200"] = _mod
_spec.loader.exec_module(_mod)
