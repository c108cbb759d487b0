"""PatchPredictor's WSI-mode host logic (coordinate grid, 255 padding outside the region) without a GPU: the patch-mode run it delegates to
is replaced by a recorder."""
from __future__ import annotations

import numpy as np
import pytest

from tiatoolbox_b200.models.architecture import IOConfig
from tiatoolbox_b200.models.engine.engine_abc import EngineABC
from tiatoolbox_b200.models.engine.patch_predictor import PatchPredictor
from tiatoolbox_b200.tools.patchextraction import PatchExtractor


def _engine():
    eng = object.__new__(PatchPredictor)
    eng.ioconfig = IOConfig(patch_input_shape=(32, 32), stride_shape=(32, 32))
    return eng


def test_wsi_mode_builds_the_reference_grid(monkeypatch):
    seen = {}

    def fake_run(self, images, *, patch_mode=True, **kwargs):
        seen["patches"], seen["kwargs"], seen["patch_mode"] = images, kwargs, patch_mode
        return {"predictions": np.arange(len(images))}

    monkeypatch.setattr(EngineABC, "run", fake_run)
    rng = np.random.default_rng(0)
    region = rng.integers(0, 255, (70, 100, 3), dtype=np.uint8)
    out = _engine().run(region, patch_mode=False, return_probabilities=True)
    coords = PatchExtractor.get_coordinates(image_shape=(100, 70), patch_input_shape=(32, 32), stride_shape=(32, 32))
    assert np.array_equal(out["coordinates"], coords) and len(coords) == 4 * 3 and seen["patch_mode"] is True
    assert seen["kwargs"] == {"return_probabilities": True}
    patches = seen["patches"]
    assert patches.shape == (12, 32, 32, 3) and patches.dtype == np.uint8
    assert np.array_equal(patches[0], region[:32, :32])
    last = patches[-1]                                        # bottom-right patch: 6 valid rows, 4 valid columns, the rest padded with 255
    assert np.array_equal(last[:6, :4], region[64:70, 96:100]) and (last[6:] == 255).all() and (last[:, 4:] == 255).all()
    # explicit patch / stride arguments override the ioconfig
    out2 = _engine().run(region, patch_mode=False, patch_input_shape=(50, 40), stride_shape=(25, 20))
    want = PatchExtractor.get_coordinates(image_shape=(100, 70), patch_input_shape=(50, 40), stride_shape=(25, 20))
    assert np.array_equal(out2["coordinates"], want) and seen["patches"].shape == (len(want), 40, 50, 3)


def test_wsi_mode_argument_errors(monkeypatch):
    monkeypatch.setattr(EngineABC, "run", lambda self, images, **k: {"predictions": np.zeros(len(images))})
    eng = _engine()
    eng.ioconfig = None
    with pytest.raises(ValueError, match="patch_input_shape"):
        eng.run(np.zeros((10, 10, 3), np.uint8), patch_mode=False)
    with pytest.raises(ValueError, match="uint8"):
        _engine().run(np.zeros((10, 10, 3), np.float32), patch_mode=False)
