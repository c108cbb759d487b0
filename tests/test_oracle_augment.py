"""StainAugmentor oracle (oracle/stain_oracle.augment) against the outputs of the unmodified reference class with fixed random draws
(tests/golden/augment_golden.npz, oracle/make_golden_augment.py), and the host-side surface of the product class."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import stain_oracle as so
from oracle.make_golden_augment import DRAWS, STAIN_MATRIX

GOLD = np.load(Path(__file__).parent / "golden" / "augment_golden.npz")


def test_oracle_equals_reference_outputs():
    img = GOLD["img"]
    for k, pairs in enumerate(DRAWS):
        for bg in (False, True):
            got = so.augment(img, STAIN_MATRIX, [p[0] for p in pairs], [p[1] for p in pairs], threshold=0.85, augment_background=bg)
            assert np.array_equal(got, GOLD[f"out{k}_{'bg' if bg else 'tissue'}"]), (k, bg)
    got = so.augment(img, GOLD["macenko_matrix"], [1.3, 0.7], [0.1, -0.15], threshold=0.85)
    assert np.array_equal(got, GOLD["macenko_out"])
    # alpha = 1, beta = 0 reproduces the image up to the least-squares residual of the 2-stain model
    assert np.abs(GOLD["out2_tissue"].astype(int) - img.astype(int)).mean() < 12


def test_product_class_surface():
    from tiatoolbox_b200.tools.stainaugment import StainAugmentor

    with pytest.raises(ValueError, match=r".*Unsupported stain extractor method.*"):
        StainAugmentor(method="invalid")
    aug = StainAugmentor(method="macenko", stain_matrix=STAIN_MATRIX, sigma1=0.3, sigma2=0.1)
    assert aug.get_transform_init_args_names() == ("method", "stain_matrix", "sigma1", "sigma2", "augment_background")
    assert aug.get_params_dependent_on_targets({}) == {}
    for _ in range(50):
        aug.get_params()
        assert 0.7 <= aug.alpha <= 1.3 and -0.1 <= aug.beta <= 0.1
    with pytest.raises(RuntimeError, match="fit"):
        aug.augment()
