"""Import shim: the package directory is `scimloperators.jl_b200/` (a dot is not importable as a module
name), so `import scimlop_b200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scimloperators.jl_b200")
_spec = importlib.util.spec_from_file_location("scimlop_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["scimlop_b200"] = _mod
_spec.loader.exec_module(_mod)
