"""Importable alias of the package directory `augmentedgaussianprocesses.jl_b200/` (its name contains a
dot and cannot be imported with a plain `import` statement):  `import agp_b200 as agp`."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "augmentedgaussianprocesses.jl_b200")
_spec = _ilu.spec_from_file_location(
    "agp_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["agp_b200"] = _mod
_spec.loader.exec_module(_mod)
