"""Import shim: loads the package directory ``3d-mask-r-cnn_b200/`` (whose name is not a
Python identifier) under the module name ``roi3d_b200``.

    import roi3d_b200
    idx = roi3d_b200.non_max_suppression_3d(boxes, scores, 1000, 0.7)
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-mask-r-cnn_b200")


def _load():
    name = __name__
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod                 # replace this shim by the real package
    spec.loader.exec_module(mod)
    return mod


_load()
