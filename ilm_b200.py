"""Import shim: the product package lives in the directory `immersedlayers.jl_b200/`
(a name Python cannot import directly because of the dot); `import ilm_b200`
loads it under this alias."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "immersedlayers.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "ilm_b200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ilm_b200"] = _mod
_spec.loader.exec_module(_mod)
