from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g

    g.build()
    from tiatoolbox_b200 import _lib

    return _lib.lib()


@pytest.fixture(scope="session")
def stain_golden():
    return np.load(ROOT / "tests" / "golden" / "stain_golden.npz")


@pytest.fixture(scope="session")
def have_reference():
    return Path("/root/reference/tiatoolbox/tools/stainnorm.py").exists()
