"""Imports the package directory `ntt-gpu-qtesla_b200/` (hyphenated, hence not importable by name)
under the module name `qtesla_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ntt-gpu-qtesla_b200")


def load():
    if "qtesla_b200" in sys.modules:
        return sys.modules["qtesla_b200"]
    spec = importlib.util.spec_from_file_location(
        "qtesla_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["qtesla_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
