"""Import alias for the package directory ``advancedps.jl_b200/``.

The directory is named after the reference repository (TuringLang/AdvancedPS.jl); a dot is not
valid in a Python module name, so ``import advancedps_b200`` loads that directory as a package.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "advancedps.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
