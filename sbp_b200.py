"""Import shim: the package directory is `summationbypartsoperatorsextra.jl_b200/` (the dot in the name makes it
unimportable by name), so `import sbp_b200` loads that directory as the package `sbp_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "summationbypartsoperatorsextra.jl_b200")
_spec = importlib.util.spec_from_file_location("sbp_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sbp_b200"] = _mod
_spec.loader.exec_module(_mod)
