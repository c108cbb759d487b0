"""GPU parity of StainAugmentor (csrc/augment.cu + the existing concentration / tissue-mask kernels) against the reference outputs with
fixed random draws.  Floating point (float32 on the device, float64 in the reference): |delta| <= 1 grey level on <= 0.5 % of the values,
and the reference's own test metric mean|delta| / 255 < 1e-2 (tests/test_stainnorm.py:148)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import stain_oracle as so
from oracle.make_golden_augment import DRAWS, STAIN_MATRIX, fixed_params

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "augment_golden.npz")


def _close(got: np.ndarray, want: np.ndarray) -> None:
    assert got.dtype == np.uint8 and got.shape == want.shape
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d > 0).mean() <= 5e-3, (d.max(), (d > 0).mean())


def test_augment_matches_reference_outputs(built_lib):
    from tiatoolbox_b200.tools.stainaugment import StainAugmentor

    img = GOLD["img"]
    for k, pairs in enumerate(DRAWS):
        for bg in (False, True):
            aug = StainAugmentor(method="macenko", stain_matrix=STAIN_MATRIX.copy(), augment_background=bg)
            aug.fit(img.copy(), threshold=0.85)
            fixed_params(aug, pairs)
            _close(aug.augment(), GOLD[f"out{k}_{'bg' if bg else 'tissue'}"])
    aug = StainAugmentor(method="macenko")
    aug.fit(img.copy())
    assert np.allclose(aug.stain_matrix, GOLD["macenko_matrix"], atol=1e-3)
    fixed_params(aug, DRAWS[0])
    got = aug.augment()
    want = so.augment(img, np.asarray(aug.stain_matrix), [1.3, 0.7], [0.1, -0.15], threshold=0.85)
    _close(got, want)


def test_tissue_mask_api_and_repeated_augment(built_lib):
    from tiatoolbox_b200.synthetic import he_tiles
    from tiatoolbox_b200.tools.stainaugment import StainAugmentor
    from tiatoolbox_b200.utils.misc import get_luminosity_tissue_mask

    img = he_tiles(1, 256, 256, seed=8)[0]
    for thr in (0.8, 0.85, 0.6):
        m = get_luminosity_tissue_mask(img, thr)
        assert m.dtype == bool and np.array_equal(m, so.get_luminosity_tissue_mask(img, thr))
    with pytest.raises(ValueError, match="Empty tissue mask"):
        get_luminosity_tissue_mask(np.full((64, 64, 3), 255, np.uint8), 0.8)
    aug = StainAugmentor(method="macenko", stain_matrix=STAIN_MATRIX.copy(), sigma1=0.4, sigma2=0.2)
    aug.fit(img)
    outs = [aug.augment() for _ in range(4)]                      # stand-alone loop of stainaugment.py:93-97: fit once, augment many
    assert all(o.shape == img.shape and o.dtype == np.uint8 for o in outs)
    assert len({o.tobytes() for o in outs}) == 4                  # fresh random draws every time
    res = aug(image=img, force_apply=True)["image"]               # albumentations-style call
    assert res.shape == img.shape and np.mean(np.abs(res / 255.0 - img / 255.0)) > 1e-3
