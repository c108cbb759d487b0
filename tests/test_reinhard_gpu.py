"""GPU parity of ReinhardNormalizer (csrc/reinhard.cu) and the device cv2-equivalent 8-bit RGB <-> Lab conversion.

Bit-exact bar: the conversions are integer code (== cv2 on all 2^24 inputs, also checked on the device below); the transformed images must
equal the outputs of the unmodified reference class (tests/golden/reinhard_golden.npz) and the oracle byte for byte."""
from __future__ import annotations

from pathlib import Path

import cv2
import numpy as np
import pytest

from oracle import reinhard_oracle as ro
from oracle.make_golden_reinhard import inputs

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "reinhard_golden.npz")


@pytest.fixture(scope="module")
def norm(built_lib):
    from tiatoolbox_b200.tools.stainnorm import ReinhardNormalizer, get_normalizer

    n = get_normalizer("reinhard")
    assert isinstance(n, ReinhardNormalizer)
    n.fit(inputs()[0])
    return n


def test_device_lab_conversion_equals_cv2_on_all_inputs(norm):
    import torch

    g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    cube = np.stack([np.stack([np.full_like(g, r), g, b], axis=-1) for r in range(256)])          # all 2^24 triples, 48 MB
    dev = torch.from_numpy(cube).cuda()
    fwd = norm.rgb2lab(dev).cpu().numpy()
    inv = norm.lab2rgb(dev).cpu().numpy()
    for r in range(0, 256):
        assert np.array_equal(fwd[r], cv2.cvtColor(cube[r], cv2.COLOR_RGB2LAB)), r
        assert np.array_equal(inv[r], cv2.cvtColor(cube[r], cv2.COLOR_LAB2RGB)), r
    small = cube[7, :5, :9]
    assert np.array_equal(norm.rgb2lab(small), cv2.cvtColor(np.ascontiguousarray(small), cv2.COLOR_RGB2LAB))   # numpy in -> numpy out


def test_fit_statistics(norm):
    np.testing.assert_allclose(norm.target_means, GOLD["target_means"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(norm.target_stds, GOLD["target_stds"], rtol=1e-12, atol=1e-13)
    assert all(isinstance(v, float) for v in (*norm.target_means, *norm.target_stds))
    m, s = norm.get_mean_std(inputs()[1][0])
    np.testing.assert_allclose(m, GOLD["tile0_means"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(s, GOLD["tile0_stds"], rtol=1e-12, atol=1e-13)
    c1, c2, c3 = norm.lab_split(inputs()[1][0])
    r1, r2, r3 = ro.lab_split(inputs()[1][0])
    assert c1.dtype == np.float32 and np.array_equal(c1, r1) and np.array_equal(c2, r2) and np.array_equal(c3, r3)
    assert np.array_equal(norm.merge_back(c1, c2, c3), ro.merge_back(r1, r2, r3))


def test_transform_equals_reference_outputs(norm):
    import torch

    target, tiles, odd, noise = inputs()
    out = norm.transform_batch(tiles)
    assert out.dtype == np.uint8 and np.array_equal(out, GOLD["tiles_out"])
    assert np.array_equal(norm.transform(odd), GOLD["odd_out"])            # H*W*3 % 4 != 0: scalar path
    assert np.array_equal(norm.transform(noise), GOLD["noise_out"])        # clip saturates
    dev_out, info = norm.transform_batch(torch.from_numpy(tiles).cuda(), return_info=True)
    assert dev_out.is_cuda and np.array_equal(dev_out.cpu().numpy(), GOLD["tiles_out"])
    np.testing.assert_allclose(info["source_means"][0].cpu().numpy(), GOLD["tile0_means"], rtol=1e-13, atol=1e-13)
    # a target given by its statistics (the reference CLI's default target image)
    norm2 = type(norm)()
    norm2.target_means, norm2.target_stds = tuple(float(v) for v in GOLD["png_means"]), tuple(float(v) for v in GOLD["png_stds"])
    norm2._fitted = True
    assert np.array_equal(norm2.transform(tiles[1]), GOLD["png_tile_out"])


def test_transform_equals_oracle_at_patch_size_and_batch(norm):
    from tiatoolbox_b200.synthetic import he_tiles

    tiles = he_tiles(20, 256, 256, seed=4)
    out = norm.transform_batch(tiles)
    for i in range(len(tiles)):
        assert np.array_equal(out[i], ro.transform(tiles[i], norm.target_means, norm.target_stds)), i
    big = he_tiles(1, 777, 1001, seed=6)[0]
    assert np.array_equal(norm.transform(big), ro.transform(big, norm.target_means, norm.target_stds))


def test_errors_like_the_reference(norm):
    from tiatoolbox_b200.tools.stainnorm import ReinhardNormalizer

    with pytest.raises(RuntimeError, match="fit"):
        ReinhardNormalizer().transform(np.zeros((8, 8, 3), np.uint8))
    flat = np.full((32, 32, 3), 180, np.uint8)
    with pytest.raises(ZeroDivisionError):
        norm.transform(flat)
    out, info = norm.transform_batch(np.stack([flat, inputs()[3][:32, :32]]), return_info=True, raise_on_error=False)
    assert info["status"].cpu().numpy().tolist() == [1, 0] and np.array_equal(out[0], flat)   # flagged image passes through
    with pytest.raises(ValueError, match="uint8"):
        norm.transform_batch(np.zeros((1, 8, 8, 3), np.float32))


def test_reinhard_as_engine_preproc(norm):
    from tests.cnn_common import resnet18_test_weights
    from tiatoolbox_b200.models.engine import PatchPredictor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = PatchPredictor("resnet18-kather100k", batch_size=4, device="cuda")
    eng.model.load_state_dict(resnet18_test_weights())
    tiles = he_tiles(5, 224, 224, seed=2)
    eng.model.preproc_func = norm.transform
    a = eng.run(tiles, return_probabilities=True)["probabilities"]
    eng.model.preproc_func = None
    b = eng.run(norm.transform_batch(tiles), return_probabilities=True)["probabilities"]
    assert np.array_equal(a, b)
