"""Import helper: the product package directory is `ikea-recommender-system_b200/` (a hyphen is not
importable), so it is loaded under the module name `ikea_recommender_system_b200`.  A symlink with
that name normally exists at the repo root; this helper also works when it does not (e.g. a copy
of the tree that dropped symlinks)."""

import importlib
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
NAME = "ikea_recommender_system_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    try:
        return importlib.import_module(NAME)
    except ModuleNotFoundError as e:
        if e.name != NAME:
            raise
    pkg_dir = os.path.join(_ROOT, "ikea-recommender-system_b200")
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
