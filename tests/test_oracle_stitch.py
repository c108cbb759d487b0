"""CPU tests of the WSI-stitching oracle (oracle/stitch_oracle.py): golden outputs of the unmodified reference
merge_batch_to_canvas, the reference's own known answers (tests/engines/test_semantic_segmentor.py:283-333), and
hand-built / property checks of the horizontal + vertical drivers."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import stitch_oracle as so

GOLDEN = Path(__file__).resolve().parent / "golden" / "stitch_golden.npz"


def grid_patches(rng, H, W, h, w, sy, sx, c=2, zero_every=0):
    """Row-major patch grid like PatchExtractor's (x fastest): returns blocks (N,h,w,c) float32 and locs (N,4) [xs,ys,xe,ye]."""
    ys = list(range(0, max(H - h, 0) + 1, sy))
    xs = list(range(0, max(W - w, 0) + 1, sx))
    if ys[-1] + h < H:
        ys.append(ys[-1] + sy)
    if xs[-1] + w < W:
        xs.append(xs[-1] + sx)
    locs = np.array([[x, y, x + w, y + h] for y in ys for x in xs], dtype=np.int64)
    blocks = rng.random((len(locs), h, w, c), dtype=np.float32) + 0.1
    if zero_every:
        blocks[::zero_every] = 0
    return blocks, locs


def test_merge_batch_to_canvas_golden():
    g = np.load(GOLDEN)
    for name in ("overlap2", "overlap4_unsorted", "disjoint_f64"):
        canvas, count = so.merge_batch_to_canvas(g[f"{name}.blocks"], g[f"{name}.locs"], tuple(g[f"{name}.shape"]))
        assert canvas.dtype == g[f"{name}.canvas"].dtype and count.dtype == np.uint8
        assert np.array_equal(canvas, g[f"{name}.canvas"]), name
        assert np.array_equal(count, g[f"{name}.count"]), name


def test_reference_known_answers():
    """tests/engines/test_semantic_segmentor.py:283-333 (non-overlapping, overlapping, zero block, empty)."""
    c, n = so.merge_batch_to_canvas(np.array([np.ones((2, 2, 1)), np.ones((2, 2, 1)) * 2]), np.array([[0, 0, 2, 2], [2, 0, 4, 2]]), (2, 4, 1))
    assert np.array_equal(c[:, :2], np.ones((2, 2, 1))) and np.array_equal(c[:, 2:], np.ones((2, 2, 1)) * 2) and np.array_equal(n, np.ones((2, 4, 1)))
    c, n = so.merge_batch_to_canvas(np.array([np.ones((2, 2, 1)), np.ones((2, 2, 1)) * 3]), np.array([[0, 0, 2, 2], [1, 0, 3, 2]]), (2, 3, 1))
    assert np.array_equal(c, np.array([[[1], [4], [3]], [[1], [4], [3]]])) and np.array_equal(n, np.array([[[1], [2], [1]], [[1], [2], [1]]]))
    c, n = so.merge_batch_to_canvas(np.array([np.zeros((2, 2, 1)), np.ones((2, 2, 1))]), np.array([[0, 0, 2, 2], [2, 0, 4, 2]]), (2, 4, 1))
    assert np.array_equal(c[:, :2], np.zeros((2, 2, 1))) and np.array_equal(n[:, :2], np.zeros((2, 2, 1))) and np.array_equal(n[:, 2:], np.ones((2, 2, 1)))
    c, n = so.merge_batch_to_canvas(np.empty((0, 2, 2, 1)), np.empty((0, 4)), (2, 2, 1))
    assert np.array_equal(c, np.zeros((2, 2, 1))) and np.array_equal(n, np.zeros((2, 2, 1), dtype=np.uint8))


def test_vertical_merge_known_answer():
    """Two row canvases of height 4 overlapping by 2 rows: seam rows are summed, counts added, result normalised."""
    a = np.full((4, 3, 1), 2.0, np.float32)
    b = np.full((4, 3, 1), 6.0, np.float32)
    ca = np.ones((4, 3, 1), np.uint8)
    cb = np.ones((4, 3, 1), np.uint8)
    cb[3] = 0   # never written -> divides by 1, stays what the canvas holds
    b[3] = 0
    locs_y = np.array([[0, 4], [0, 4], [2, 6], [2, 6]])
    out = so.merge_vertical_chunkwise([a, b], [ca, cb], locs_y)
    assert out.shape == (6, 3, 1)
    assert np.array_equal(out[:2], np.full((2, 3, 1), 2.0)) and np.array_equal(out[2:4], np.full((2, 3, 1), 4.0))
    assert np.array_equal(out[4], np.full((3, 1), 6.0)) and np.array_equal(out[5], np.zeros((3, 1)))
    clipped = so.merge_vertical_chunkwise([a, b], [ca, cb], locs_y, output_shape=(5, 2))
    assert clipped.shape == (5, 2, 1) and np.array_equal(clipped, out[:5, :2])


@pytest.mark.parametrize("cfg", [(40, 52, 16, 16, 12, 12, 0), (37, 45, 16, 20, 8, 15, 3), (16, 16, 16, 16, 16, 16, 0)])
def test_stitch_equals_global_overlap_add(cfg):
    """With at most two patch rows per pixel the chunkwise pipeline is a global overlap-add / count (float association: per row first)."""
    H, W, h, w, sy, sx, zero_every = cfg
    rng = np.random.default_rng(7)
    blocks, locs = grid_patches(rng, H, W, h, w, sy, sx, zero_every=zero_every)
    Wc = int(locs[:, 2].max())
    out = so.stitch(blocks, locs, canvas_width=Wc, output_shape=(H, W))
    Hc = int(locs[:, 3].max())
    acc = np.zeros((Hc, Wc, blocks.shape[-1]), np.float64)
    cnt = np.zeros((Hc, Wc, 1), np.int64)
    for b, (xs, ys, xe, ye) in zip(blocks, locs):
        if not b.any():
            continue
        acc[ys:ye, xs:xe] += b
        cnt[ys:ye, xs:xe] += 1
    ref = (acc / np.where(cnt == 0, 1, cnt))[:H, :W]
    assert out.shape == ref.shape and out.dtype == np.float32
    assert np.allclose(out, ref, rtol=2e-6, atol=1e-7)
