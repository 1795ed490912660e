"""mgx_loader.py — imports the package directory ``mygram-db_b200`` (hyphenated, hence not a
valid module name) under the module name ``mygram_db_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
_NAME = "mygram_db_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    path = os.path.join(ROOT, "mygram-db_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(_NAME, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
