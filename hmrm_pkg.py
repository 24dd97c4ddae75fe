"""Import helper: the package directory is named `heightmap-ray-marcher_b200` (a hyphen is
not importable), so load it under the module name `heightmap_ray_marcher_b200`."""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent / "heightmap-ray-marcher_b200"
MODULE_NAME = "heightmap_ray_marcher_b200"


def load():
    if MODULE_NAME in sys.modules:
        return sys.modules[MODULE_NAME]
    spec = importlib.util.spec_from_file_location(
        MODULE_NAME, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[MODULE_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
