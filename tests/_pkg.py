"""Load the hyphenated package directory openfoam-dev_b200/ under the importable alias `b200ls`."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def load_pkg():
    if "b200ls" in sys.modules:
        return sys.modules["b200ls"]
    pkg = ROOT / "openfoam-dev_b200"
    spec = importlib.util.spec_from_file_location("b200ls", pkg / "__init__.py",
                                                  submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b200ls"] = mod
    spec.loader.exec_module(mod)
    return mod
