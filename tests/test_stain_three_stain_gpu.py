"""`StainNormalizer.get_concentrations` with a 3x3 user stain matrix on the device.  The same tensor-op function is checked
against the unmodified reference on the CPU device in tests/test_oracle_stain.py; this module repeats it on cuda:0.  Added after the last GPU
call of round 2, so it sorts last in tests/conftest.py."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_three_stain_concentrations_on_the_device(built_lib):  # noqa: ARG001
    from oracle import stain_oracle as so
    from tiatoolbox_b200.synthetic import he_tiles
    from tiatoolbox_b200.tools import stainnorm

    t = he_tiles(1, 64, 64, seed=9)[0]
    m = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11], [0.27, 0.57, 0.78]])
    got = stainnorm.StainNormalizer.get_concentrations(t, m)
    od = so.rgb2od(t.copy()).reshape(-1, 3)
    ref = np.linalg.lstsq(m.T, od.T, rcond=-1)[0].T
    assert got.shape == ref.shape == (64 * 64, 3)
    np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9)
    with pytest.raises(ValueError, match="shape"):
        stainnorm.StainNormalizer.get_concentrations(t, np.ones((4, 3)))
