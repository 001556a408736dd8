"""Import shim: the product package lives in `laplace-inducing-points_b200/` (a directory name that is not a
valid Python identifier), so `import lip_b200` registers that directory as the package `lip_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "laplace-inducing-points_b200")
_spec = importlib.util.spec_from_file_location("lip_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lip_b200"] = _mod
_spec.loader.exec_module(_mod)
