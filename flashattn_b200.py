"""Import shim: the package directory is named ``llmsys-project-flashattn_b200`` (not a valid
Python identifier), so ``import flashattn_b200`` loads it from there under this name."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "llmsys-project-flashattn_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
