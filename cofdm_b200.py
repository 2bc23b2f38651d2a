"""Import shim: the package directory is `c-ofdm_b200/` (named after the reference, not a valid Python
identifier); this module loads it under the importable name `cofdm_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c-ofdm_b200")
_spec = importlib.util.spec_from_file_location("cofdm_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cofdm_b200"] = _mod
_spec.loader.exec_module(_mod)
