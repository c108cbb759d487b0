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


# The long-established hot-path modules run first, the newest modules last, so that with `-x` a failure in the newest code cannot hide the
# results of the hot-path tests.  Every module listed here except the last has passed on a B200 (profiles/r2t_gpu_tests.log: 117 passed); the last one was
# added after the final GPU call of round 2 (same function CPU-checked in tests/test_oracle_stain.py).
_NEWEST_GPU_MODULES = ("test_reinhard_gpu", "test_augment_gpu", "test_contours_gpu", "test_wsi_gpu", "test_hovernetplus_gpu",
                       "test_reinhard_fused_gpu", "test_watershed_v2_gpu", "test_net_capi_gpu", "test_stain_three_stain_gpu")


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
