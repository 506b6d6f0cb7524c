"""The package directory is named `globalsensitivity.jl_b200` (not a valid Python identifier), so it is imported
under the module name `gsa_b200`:

    import gsa_loader; gsa_b200 = gsa_loader.load()
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "globalsensitivity.jl_b200")


def load():
    if "gsa_b200" in sys.modules:
        return sys.modules["gsa_b200"]
    spec = importlib.util.spec_from_file_location("gsa_b200", os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gsa_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
