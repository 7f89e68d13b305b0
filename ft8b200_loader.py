"""Import helper: the package directory is named ``rtlsdr-ft8d_b200`` (hyphen, after the reference), which
Python cannot import by name; load() registers it as ``rtlsdr_ft8d_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    name = "rtlsdr_ft8d_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "rtlsdr-ft8d_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
