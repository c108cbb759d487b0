"""GPU parity tests of the stain path: CUDA (through the C ABI / the Python drop-in classes)
versus the numpy oracle and the committed reference fixtures.

Tolerances (stated per north_star: 1e-3 relative on floats): stain matrices / concentrations
percentiles are float work -> atol/rtol 1e-3 or tighter; the uint8 image is the float result
truncated, so a float error of ~1e-5 flips a small fraction of pixels by one level:
we assert |delta| <= 1 everywhere, <= 2 % flipped, and the reference's own metric
mean|delta|/255 < 1e-2 (/root/reference/tests/test_stainnorm.py:148).  The tissue mask is integer
work and must be bit-exact.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
import pytest

from oracle import stain_oracle as so
from tiatoolbox_b200.synthetic import he_target, he_tiles

pytestmark = pytest.mark.gpu

CUSTOM = np.array([[0.60, 0.72, 0.34], [0.10, 0.95, 0.30]])


@pytest.fixture(scope="module")
def api(built_lib):
    from tiatoolbox_b200 import _lib
    from tiatoolbox_b200.tools import stainnorm

    return _lib, stainnorm


def _norm(stainnorm, method):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if method == "custom":
            return stainnorm.get_normalizer("custom", stain_matrix=CUSTOM)
        return stainnorm.get_normalizer(method)


def _check_u8(out, ref):
    d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1, f"max |delta| {d.max()}"
    assert (d > 0).mean() <= 0.02, f"flipped fraction {(d > 0).mean()}"
    assert d.mean() / 255.0 < 1e-2


def _set_path(_lib, path):
    assert _lib.lib().tiab_stain_set_path(_lib.handle(0), path) == 0


@pytest.mark.parametrize("method", ["macenko", "vahadane", "ruifrok", "custom"])
def test_fit_matches_oracle_and_reference_fixture(api, stain_golden, method):
    _lib, stainnorm = api
    _set_path(_lib, 0)
    target = he_target(512, 512, seed=1234)  # 786 KB: general (multi-pass) path
    n = _norm(stainnorm, method)
    n.fit(target)
    m_t, max_c = so.fit(method, target.copy(), CUSTOM if method == "custom" else None)
    assert np.allclose(n.stain_matrix_target, m_t, atol=2e-4)
    assert np.allclose(n.maxC_target, max_c, rtol=5e-4)
    # the unmodified reference's answer for the same target (tests/golden, oracle/make_golden.py)
    assert np.allclose(n.stain_matrix_target, stain_golden[f"{method}_synfit_M"], atol=2e-4)
    assert np.allclose(n.maxC_target, stain_golden[f"{method}_synfit_maxC"], rtol=5e-4)
    assert n.stain_matrix_target_RGB.shape == (2, 3)


@pytest.mark.parametrize("method", ["macenko", "vahadane", "ruifrok", "custom"])
def test_fit_on_the_canonical_target_matches_the_reference(api, stain_golden, method):
    """`fit` on the reference's in-tree 1000 x 1000 target (tiatoolbox/data/target_image.png, stainnorm.py:68-87) against the UNMODIFIED
    reference's own answers (tests/golden/stain_golden.npz `*_realfit_*`, oracle/make_golden.py:55-57).  This is the one input where
    Vahadane's > 8.4e5-tissue-pixel LARS early-stop quirk bites.  The image is read from /root/reference or from the
    vendored checker copy oracle/_ref (oracle/vendor_ref.py: it travels to the GPU box, it is never committed)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference data file not present (neither /root/reference nor oracle/_ref)")
    _lib, stainnorm = api
    _set_path(_lib, 0)
    target = ref_shim.target_image()
    assert target.shape == (1000, 1000, 3)
    n = _norm(stainnorm, method)
    n.fit(target)
    want_m, want_c = stain_golden[f"{method}_realfit_M"], stain_golden[f"{method}_realfit_maxC"]
    err_m = float(np.abs(n.stain_matrix_target - want_m).max())
    err_c = float(np.abs(n.maxC_target / want_c - 1).max())
    print(f"\n{method}: fit on target_image.png  max|dM| = {err_m:.2e}  max rel dmaxC = {err_c:.2e}")
    assert err_m <= 5e-4 and err_c <= 1e-3, (err_m, err_c)


@pytest.mark.parametrize("path", [1, 2], ids=["fused", "general"])
@pytest.mark.parametrize("method", ["macenko", "vahadane", "ruifrok", "custom"])
def test_transform_matches_oracle(api, stain_golden, method, path):
    _lib, stainnorm = api
    _set_path(_lib, path)
    try:
        n = _norm(stainnorm, method)
        # use the reference's own fit state so that this test isolates transform()
        n.stain_matrix_target = stain_golden[f"{method}_synfit_M"]
        n.maxC_target = stain_golden[f"{method}_synfit_maxC"]
        tiles = he_tiles(6, 256, 256, seed=0)
        out, info = n.transform_batch(tiles, return_info=True)
        assert out.shape == tiles.shape and out.dtype == np.uint8
        assert (info["status"] == 0).all()
        sm = CUSTOM if method == "custom" else None
        for i, t in enumerate(tiles):
            dbg = {}
            ref = so.transform_float(method, t.copy(), n.stain_matrix_target, n.maxC_target, sm, debug=dbg)
            assert np.allclose(info["stain_matrix_source"][i], dbg["stain_matrix_source"], atol=5e-4), i
            assert np.allclose(info["maxC_source"][i], dbg["max_c_source"].ravel(), rtol=1e-3), i
            _check_u8(out[i], ref.astype(np.uint8))
        if method in ("macenko", "vahadane"):  # the reference's own outputs (fixture)
            for i in range(2):
                _check_u8(out[i], stain_golden[f"{method}_t256_out"][i])
                assert np.allclose(info["stain_matrix_source"][i], stain_golden[f"{method}_t256_Msrc"][i], atol=5e-4)
    finally:
        _set_path(_lib, 0)


@pytest.mark.parametrize("method", ["macenko", "ruifrok", "vahadane"])
def test_small_and_ragged_shapes(api, stain_golden, method):
    _lib, stainnorm = api
    n = _norm(stainnorm, method)
    n.stain_matrix_target = stain_golden[f"{method}_synfit_M"]
    n.maxC_target = stain_golden[f"{method}_synfit_maxC"]
    for name, tiles in (("t64", he_tiles(4, 64, 64, seed=7)), ("todd", he_tiles(2, 100, 75, seed=11))):
        for i, t in enumerate(tiles):
            out = n.transform(t)  # single-image reference API
            assert out.shape == t.shape
            _check_u8(out, stain_golden[f"{method}_{name}_out"][i])
    odd = he_tiles(1, 75, 75, seed=5)[0]  # H*W % 4 != 0: byte-wise pixel path
    ref = so.transform(method, odd.copy(), n.stain_matrix_target, n.maxC_target)
    _check_u8(n.transform(odd), ref)


def test_float_output_within_1e3(api, stain_golden):
    """Pre-quantisation float parity (rel 1e-3 of full scale) through the raw C ABI, both paths."""
    import torch

    _lib, stainnorm = api
    L = _lib.lib()
    h = _lib.handle(0)
    tiles = he_tiles(3, 256, 256, seed=0)
    n = _norm(stainnorm, "macenko")
    mt = np.ascontiguousarray(stain_golden["macenko_synfit_M"]).reshape(-1)
    mc = np.ascontiguousarray(stain_golden["macenko_synfit_maxC"]).reshape(-1)
    p = n.extractor._params()
    x = torch.from_numpy(tiles).cuda()
    for path in (1, 2):
        _set_path(_lib, path)
        out = torch.empty_like(x)
        f32 = torch.empty(x.shape, dtype=torch.float32, device="cuda")
        rc = L.tiab_stain_transform_batch(h, C.byref(p), x.data_ptr(), out.data_ptr(), 3, 256, 256, _lib.dptr(mt),
                                          _lib.dptr(mc), None, None, None, f32.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
        _lib.check(h, rc)
        got = f32.cpu().numpy()
        for i in range(3):
            ref = so.transform_float("macenko", tiles[i].copy(), stain_golden["macenko_synfit_M"], stain_golden["macenko_synfit_maxC"])
            assert np.abs(got[i] - ref).max() <= 1e-3 * 255.0
            assert np.array_equal(out[i].cpu().numpy(), np.floor(got[i]).astype(np.uint8))
    _set_path(_lib, 0)


def test_tissue_mask_bit_exact(api, stain_golden):
    import torch

    _lib, stainnorm = api
    L = _lib.lib()
    h = _lib.handle(0)
    n = _norm(stainnorm, "macenko")
    p = n.extractor._params()
    assert p.y_threshold == 1146
    tiles = np.concatenate([he_tiles(3, 256, 256, seed=0), np.random.default_rng(5).integers(0, 256, (2, 256, 256, 3), dtype=np.uint8)])
    x = torch.from_numpy(tiles).cuda()
    mask = torch.empty(tiles.shape[:3], dtype=torch.uint8, device="cuda")
    status = torch.empty(len(tiles), dtype=torch.int32, device="cuda")
    rc = L.tiab_stain_tissue_mask(h, C.byref(p), x.data_ptr(), mask.data_ptr(), len(tiles), 256, 256, status.data_ptr(), None)
    _lib.check(h, rc)
    torch.cuda.synchronize()
    got = mask.cpu().numpy().astype(bool)
    for i, t in enumerate(tiles):
        assert np.array_equal(got[i], so.get_luminosity_tissue_mask(t, 0.8)), i
    assert np.array_equal(got[:2], stain_golden["t256_mask"])


def test_empty_tissue_mask_raises(api, stain_golden):
    _lib, stainnorm = api
    n = _norm(stainnorm, "macenko")
    white = np.full((256, 256, 3), 255, np.uint8)
    with pytest.raises(ValueError, match="Empty tissue mask"):
        n.fit(white)
    n.stain_matrix_target = stain_golden["macenko_synfit_M"]
    n.maxC_target = stain_golden["macenko_synfit_maxC"]
    with pytest.raises(ValueError, match="Empty tissue mask"):
        n.transform(white)
    with pytest.raises(ValueError, match="Empty tissue mask"):
        n.extractor.get_stain_matrix(white)
    # batch: the bad image is passed through unchanged and flagged, the others are normalised
    tiles = he_tiles(2, 256, 256, seed=0)
    batch = np.stack([tiles[0], white, tiles[1]])
    out, info = n.transform_batch(batch, return_info=True, raise_on_error=False)
    assert list(info["status"]) == [0, 1, 0]
    assert np.array_equal(out[1], white)
    _check_u8(out[2], stain_golden["macenko_t256_out"][1])


def test_api_errors(api):
    _lib, stainnorm = api
    from tiatoolbox_b200.utils.exceptions import MethodNotSupportedError

    with pytest.raises(MethodNotSupportedError):
        stainnorm.get_normalizer("invalid")
    with pytest.raises(ValueError, match="`stain_matrix` is only defined when using"):
        stainnorm.get_normalizer("ruifrok", stain_matrix=CUSTOM)
    with pytest.raises(ValueError, match="`stain_matrix` is None when using"):
        stainnorm.get_normalizer("custom")
    with pytest.raises(ValueError, match="Stain matrix must have shape"):
        stainnorm.CustomNormalizer(np.zeros((4, 3)))


def test_cuda_tensor_in_place_and_batch_invariance(api, stain_golden):
    import torch

    _lib, stainnorm = api
    n = _norm(stainnorm, "macenko")
    n.stain_matrix_target = stain_golden["macenko_synfit_M"]
    n.maxC_target = stain_golden["macenko_synfit_maxC"]
    tiles = he_tiles(300, 256, 256, seed=0)  # > 148 CTAs: exercises the persistent loop
    x = torch.from_numpy(tiles).cuda()
    out = n.transform_batch(x)
    assert out.is_cuda and out.shape == x.shape
    first = n.transform_batch(tiles[:2])
    assert np.array_equal(out[:2].cpu().numpy(), first)  # result independent of batch composition
    perm = torch.randperm(300, generator=torch.Generator().manual_seed(0))
    out_p = n.transform_batch(x[perm.cuda()].contiguous())
    assert torch.equal(out_p, out[perm.cuda()])
    _check_u8(out[1].cpu().numpy(), stain_golden["macenko_t256_out"][1])


def test_get_concentrations(api):
    _lib, stainnorm = api
    t = he_tiles(1, 64, 64, seed=9)[0]
    got = stainnorm.StainNormalizer.get_concentrations(t, so.RUIFROK)
    ref = so.get_concentrations(t.copy(), so.RUIFROK)
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5)
