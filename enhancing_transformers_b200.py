"""Import shim: the package directory is named ``enhancing-transformers_b200`` (a hyphen is not a
valid module name), so ``import enhancing_transformers_b200`` loads it from that directory and
registers it under this importable name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "enhancing-transformers_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
