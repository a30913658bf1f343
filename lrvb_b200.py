"""Import shim: loads ``linearresponsevariationalbayes.py_b200/`` (a directory name Python cannot
import directly because of the dot) as the package ``lrvb_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "linearresponsevariationalbayes.py_b200")
_spec = importlib.util.spec_from_file_location(
    "lrvb_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lrvb_b200"] = _mod
_spec.loader.exec_module(_mod)
