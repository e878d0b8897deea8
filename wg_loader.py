"""Import helper: the package directory is named `cuda-winograd_b200` (not a Python identifier), so it is loaded by
path under the module name `cuda_winograd_b200`. Used by tests/, bench.py and __graft_entry__.py."""
import importlib.util
import os
import sys

_NAME = "cuda_winograd_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-winograd_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
