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


# GPU test modules added last in round 1 (each has passed on a B200 once: profiles/r1h_gpu_tests_*.log): they are collected AFTER the
# long-established modules, so that with `-x` a failure in the newest code cannot hide the results of the hot-path tests.
_NEWEST_GPU_MODULES = ("test_reinhard_gpu", "test_augment_gpu", "test_contours_gpu", "test_wsi_gpu", "test_hovernetplus_gpu",
                       "test_watershed_v2_gpu")   # the last one has NOT run on a B200 yet (opt-in code path, written after the budget was spent)


def pytest_collection_modifyitems(config, items):  # noqa: ARG001
    def late(item) -> int:
        name = Path(str(item.fspath)).stem
        return _NEWEST_GPU_MODULES.index(name) + 1 if name in _NEWEST_GPU_MODULES else 0

    items.sort(key=late)   # stable: everything else keeps its order


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
