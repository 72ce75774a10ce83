"""Import helper: the package directory is named `climateparameterizations.jl_b200` (a dot is not importable
as a plain module name), so it is loaded under the alias `cpz_b200`."""
import importlib.util
import os
import sys

ALIAS = "cpz_b200"


def load():
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "climateparameterizations.jl_b200")
    spec = importlib.util.spec_from_file_location(ALIAS, os.path.join(root, "__init__.py"),
                                                  submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
