"""Import helper: the package directory is named ``gnss-gps-sdr_b200`` (with a hyphen, as the
build contract asks), which ``import`` cannot spell.  ``load()`` registers it as module
``gnss_gps_sdr_b200``."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "gnss-gps-sdr_b200"
PKG_NAME = "gnss_gps_sdr_b200"


def load():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(PKG_NAME, PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
