"""CPU tests: pin the numpy oracle (oracle/stain_oracle.py) against the reference's golden vector,
the committed fixtures generated from the unmodified reference, and (when /root/reference is
present) the reference itself through the source shim."""
from __future__ import annotations

import warnings

import numpy as np
import pytest

from oracle import stain_oracle as so
from tiatoolbox_b200.synthetic import he_target, he_tiles


def test_contrast_enhancer_reference_golden_vector(stain_golden):
    # /root/reference/tests/test_utils.py:882-911
    out = so.contrast_enhancer(stain_golden["contrast_in"], 2, 98)
    assert np.array_equal(out, stain_golden["contrast_out"])
    with pytest.raises(AssertionError):
        so.contrast_enhancer(np.float32(stain_golden["contrast_in"]), 2, 98)


def test_lab_l_matches_cv2_exhaustively():
    cv2 = pytest.importorskip("cv2")
    # all 2^24 colours, in 16 slabs
    g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for r0 in range(0, 256, 16):
        r = np.arange(r0, r0 + 16, dtype=np.uint8)
        img = np.empty((16, 256, 256, 3), np.uint8)
        img[..., 0] = r[:, None, None]
        img[..., 1] = g[None]
        img[..., 2] = b[None]
        img = img.reshape(16 * 256, 256, 3)
        ref = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)[..., 0]
        assert np.array_equal(so.lab_l(img), ref)
    assert so.y_threshold(0.8) == 1146


def test_tissue_mask_golden(stain_golden):
    tiles = he_tiles(2, 256, 256, seed=0)
    for i, t in enumerate(tiles):
        assert np.array_equal(so.get_luminosity_tissue_mask(t, 0.8), stain_golden["t256_mask"][i])
    with pytest.raises(ValueError, match="Empty tissue mask"):
        so.get_luminosity_tissue_mask(np.zeros((100, 100, 3)), threshold=0)


@pytest.mark.parametrize("method", ["macenko", "ruifrok", "vahadane", "custom"])
def test_oracle_matches_reference_fixtures(stain_golden, method):
    g = stain_golden
    sm = g["custom_matrix"] if method == "custom" else None
    target = he_target(512, 512, seed=1234)
    m_t, max_c = so.fit(method, target.copy(), sm)
    assert np.allclose(m_t, g[f"{method}_synfit_M"], atol=1e-10)
    assert np.allclose(max_c, g[f"{method}_synfit_maxC"], atol=1e-9)
    sets = {"t64": he_tiles(4, 64, 64, seed=7), "todd": he_tiles(2, 100, 75, seed=11)}
    if method in ("macenko", "vahadane"):
        sets["t256"] = he_tiles(2, 256, 256, seed=0)
    for name, tiles in sets.items():
        for i, t in enumerate(tiles):
            dbg = {}
            out = so.transform_float(method, t.copy(), g[f"{method}_synfit_M"], g[f"{method}_synfit_maxC"], sm, debug=dbg)
            assert np.allclose(dbg["stain_matrix_source"], g[f"{method}_{name}_Msrc"][i], atol=1e-9)
            assert np.array_equal(out.astype(np.uint8), g[f"{method}_{name}_out"][i])


def test_oracle_matches_live_reference(have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import ref_shim

    sn = ref_shim.load_stain()
    target = ref_shim.target_image()
    tiles = he_tiles(2, 128, 128, seed=21)
    for method in ("macenko", "ruifrok", "vahadane"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            n = sn.get_normalizer(method)
        n.fit(target.copy())
        m_t, max_c = so.fit(method, target.copy())
        assert np.allclose(m_t, n.stain_matrix_target, atol=1e-10), method
        assert np.allclose(max_c, n.maxC_target, atol=1e-9), method
        for t in tiles:
            assert np.array_equal(so.transform(method, t.copy(), m_t, max_c), n.transform(t.copy()))


def test_lasso2_matches_sklearn_lars():
    lars = pytest.importorskip("sklearn.linear_model._least_angle")
    rng = np.random.default_rng(0)
    for t in range(400):
        n = int(10 ** rng.uniform(1, 6.5))
        d = rng.normal(size=(2, 50)) * rng.uniform(0.1, 3)
        if t % 3 == 0:
            d[1] = d[0] * rng.uniform(0.5, 1.5) + rng.normal(size=50) * 0.05
        y = rng.normal(size=50) * rng.uniform(0.01, 3)
        gram, c = d @ d.T, d @ y
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, _, coefs, _ = lars._lars_path_solver(
                X=None, y=None, Xy=c.copy(), Gram=gram.copy(), n_samples=n, alpha_min=0.1 / n,
                method="lasso", return_path=False, return_n_iter=True, max_iter=1000,
            )
        mine = so.lasso2(gram, c, 0.1, n)
        assert np.allclose(mine, coefs, rtol=1e-8, atol=1e-10)


def test_three_stain_concentrations_match_the_reference_lstsq():
    """`StainNormalizer.get_concentrations` with a 3x3 user matrix (the only shape besides 2x3 that CustomExtractor accepts): the tensor-op
    spelling used on the device equals the unmodified reference's lstsq on the same pixels (run on the CPU device here)."""
    import torch

    from oracle import ref_shim
    from tiatoolbox_b200.tools import stainnorm

    if not ref_shim.available():
        pytest.skip("reference sources not present")
    sn = ref_shim.load_stain()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 41, 3), dtype=np.uint8)
    img[0, :5] = 0
    m = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11], [0.27, 0.57, 0.78]])
    ref = sn.StainNormalizer.get_concentrations(img.copy(), m)
    got = stainnorm._concentrations_three_stains(torch.from_numpy(img), m).numpy()
    assert got.shape == ref.shape == (37 * 41, 3)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-10)
