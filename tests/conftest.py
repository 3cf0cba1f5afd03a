import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the C-ABI library and the C oracle once
    root = Path(__file__).resolve().parent.parent
    if not (root / "openfoam-dev_b200" / "libb200ls.so").exists() or not (root / "oracle" / "libldu_oracle.so").exists():
        from _pkg import load_pkg

        load_pkg()
        from b200ls import build

        build.build_lib()
        build.build_oracle()


@pytest.fixture(scope="session")
def pkg():
    from _pkg import load_pkg

    return load_pkg()
