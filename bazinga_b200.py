"""Import shim: the package directory is named ``bazinga.jl_b200`` (after aldma/Bazinga.jl), which is
not a valid Python identifier, so this module loads it under the importable name ``bazinga_b200``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bazinga.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "bazinga_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bazinga_b200"] = _mod
_spec.loader.exec_module(_mod)
