"""Import shim: the package directory is named ``fourier-neural-mappings_b200`` (not a valid Python
identifier), so ``import fnm_b200`` loads it under this name.  After the first import
``fnm_b200`` *is* the package (sub-modules import as ``fnm_b200.shared`` etc.)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fourier-neural-mappings_b200")
_spec = importlib.util.spec_from_file_location(
    "fnm_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fnm_b200"] = _mod
_spec.loader.exec_module(_mod)
