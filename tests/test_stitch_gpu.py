"""GPU parity of the device WSI stitcher (csrc/stitch.cu) against the oracle restatement of the reference's merge functions
(bit-exact float32), the committed outputs of the unmodified reference, and the SemanticSegmentor region mode."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import stitch_oracle as so
from tests.test_oracle_stitch import grid_patches

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden" / "stitch_golden.npz"


def test_merge_batch_to_canvas_matches_reference_golden(built_lib):
    from tiatoolbox_b200.models.engine.semantic_segmentor import merge_batch_to_canvas

    g = np.load(GOLDEN)
    for name in ("overlap2", "overlap4_unsorted"):
        canvas, count = merge_batch_to_canvas(g[f"{name}.blocks"], g[f"{name}.locs"], tuple(g[f"{name}.shape"]))
        assert canvas.dtype == np.float32 and count.dtype == np.uint8 and count.shape == (*canvas.shape[:2], 1)
        assert np.array_equal(canvas, g[f"{name}.canvas"]), name
        assert np.array_equal(count, g[f"{name}.count"]), name
    with pytest.raises(TypeError):
        merge_batch_to_canvas(g["disjoint_f64.blocks"], g["disjoint_f64.locs"], tuple(g["disjoint_f64.shape"]))


def test_reference_known_answers_on_device(built_lib):
    """tests/engines/test_semantic_segmentor.py:283-333 with float32 blocks."""
    from tiatoolbox_b200.models.engine.semantic_segmentor import merge_batch_to_canvas

    f = np.float32
    c, n = merge_batch_to_canvas(np.array([np.ones((2, 2, 1), f), np.ones((2, 2, 1), f) * 2]), np.array([[0, 0, 2, 2], [2, 0, 4, 2]]), (2, 4, 1))
    assert np.array_equal(c[:, :2], np.ones((2, 2, 1))) and np.array_equal(c[:, 2:], np.ones((2, 2, 1)) * 2) and np.array_equal(n, np.ones((2, 4, 1)))
    c, n = merge_batch_to_canvas(np.array([np.ones((2, 2, 1), f), np.ones((2, 2, 1), f) * 3]), np.array([[0, 0, 2, 2], [1, 0, 3, 2]]), (2, 3, 1))
    assert np.array_equal(c, np.array([[[1], [4], [3]], [[1], [4], [3]]])) and np.array_equal(n, np.array([[[1], [2], [1]], [[1], [2], [1]]]))
    c, n = merge_batch_to_canvas(np.array([np.zeros((2, 2, 1), f), np.ones((2, 2, 1), f)]), np.array([[0, 0, 2, 2], [2, 0, 4, 2]]), (2, 4, 1))
    assert np.array_equal(c[:, :2], np.zeros((2, 2, 1))) and np.array_equal(n[:, :2], np.zeros((2, 2, 1))) and np.array_equal(n[:, 2:], np.ones((2, 2, 1)))
    c, n = merge_batch_to_canvas(np.empty((0, 2, 2, 1), f), np.empty((0, 4)), (2, 2, 1))
    assert np.array_equal(c, np.zeros((2, 2, 1))) and np.array_equal(n, np.zeros((2, 2, 1), dtype=np.uint8))


@pytest.mark.parametrize("cfg", [(40, 52, 16, 16, 12, 12, 0), (37, 45, 16, 20, 8, 15, 3), (16, 16, 16, 16, 16, 16, 0), (300, 517, 64, 64, 32, 48, 7)])
def test_stitch_patches_bit_exact_vs_oracle(built_lib, cfg):
    import torch

    from tiatoolbox_b200.models.engine.semantic_segmentor import stitch_patches

    H, W, h, w, sy, sx, zero_every = cfg
    rng = np.random.default_rng(11)
    blocks, locs = grid_patches(rng, H, W, h, w, sy, sx, c=3, zero_every=zero_every)
    Wc = int(locs[:, 2].max())
    ref = so.stitch(blocks, locs, canvas_width=Wc, output_shape=(H, W))
    got = stitch_patches(blocks, locs, canvas_width=Wc, output_shape=(H, W))
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    dev = stitch_patches(torch.from_numpy(blocks).cuda(), locs, canvas_width=Wc, output_shape=(H, W))   # CUDA in -> CUDA out
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), ref)


def test_stitch_full_size_property(built_lib):
    """2048 x 2048 region, 512-pixel tiles at stride 256 (every interior pixel covered by 4 tiles): constant tiles must stitch to the
    constant, counts never leak, and the map equals the oracle on a strip."""
    import torch

    from tiatoolbox_b200.models.engine.semantic_segmentor import patch_grid, stitch_patches

    locs = patch_grid((2048, 2048), (512, 512), (256, 256))
    blocks = torch.full((len(locs), 512, 512, 2), 0.25, device="cuda")
    blocks[..., 1] = 0.75
    out = stitch_patches(blocks, locs, canvas_width=2048, output_shape=(2048, 2048))
    assert out.shape == (2048, 2048, 2)
    assert torch.equal(out[..., 0], torch.full((2048, 2048), 0.25, device="cuda")) and torch.equal(out[..., 1], torch.full((2048, 2048), 0.75, device="cuda"))


def test_stitch_rejects_three_row_overlap(built_lib):
    from tiatoolbox_b200._lib import TiabError
    from tiatoolbox_b200.models.engine.semantic_segmentor import patch_grid, stitch_patches

    locs = patch_grid((64, 32), (32, 32), (8, 32))          # stride < h/2: three patch rows cover a pixel
    with pytest.raises(TiabError):
        stitch_patches(np.ones((len(locs), 32, 32, 1), np.float32), locs, canvas_width=32, output_shape=(64, 32))


def test_semantic_segmentor_region_mode(built_lib):
    """run(image, patch_mode=False): stitched probabilities == oracle stitching of the per-tile device outputs."""
    import torch

    from tests.cnn_common import unet_test_weights
    from tiatoolbox_b200.models.engine import SemanticSegmentor
    from tiatoolbox_b200.models.engine.semantic_segmentor import patch_grid
    from tiatoolbox_b200.synthetic import he_tiles

    eng = SemanticSegmentor("fcn-tissue_mask", batch_size=4, device="cuda")
    eng.model.load_state_dict(unet_test_weights())
    region = np.concatenate([np.concatenate(list(he_tiles(2, 256, 256, seed=s)), axis=1) for s in (0, 1)], axis=0)[:448, :480]   # (448, 480, 3)
    out = eng.run(region, patch_mode=False, patch_input_shape=(256, 256), stride_shape=(128, 128), return_probabilities=True)
    assert out["predictions"].shape == (224, 240) and out["probabilities"].shape == (224, 240, 2) and out["predictions"].dtype == bool
    # rebuild from the per-tile outputs with the oracle
    locs_in = patch_grid(region.shape[:2], (256, 256), (128, 128))
    padded = np.zeros((int(locs_in[:, 3].max()), int(locs_in[:, 2].max()), 3), np.uint8)
    padded[:448, :480] = region
    tiles = np.stack([padded[y0:y1, x0:x1] for x0, y0, x1, y1 in locs_in])
    per_tile = eng.run(tiles, patch_mode=True, return_probabilities=True)["probabilities"]
    ref = so.stitch(per_tile, locs_in // 2, canvas_width=int(locs_in[:, 2].max()) // 2, output_shape=(224, 240))
    assert np.array_equal(out["probabilities"], ref)
    assert np.array_equal(out["predictions"].astype(np.uint8), ref.argmax(-1).astype(np.uint8))
    torch.cuda.synchronize()
