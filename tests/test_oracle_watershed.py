"""The version-2 flood (csrc/watershed_math.cuh: 8-ary heap, total order (value, age, index)) compiled for the host, against the oracle's
restatement of skimage.segmentation.watershed — on the pipeline's own distance maps / markers (synthetic nuclei) and on adversarial inputs
full of exact ties (plateaus, quantised values), where the shape of a priority queue would show if the order were not total."""
from __future__ import annotations

import numpy as np

from oracle import hovernet_postproc_oracle as po
from tiatoolbox_b200.synthetic import hovernet_head_maps


def _flood(lib, dist: np.ndarray, markers: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Both host-compiled data flows must agree: the whole-image flood (version 2's code) and the per-component box-local flood (version 3's)."""
    d = np.ascontiguousarray(dist, dtype=np.float64)
    m = np.ascontiguousarray(markers, dtype=np.int32)
    k = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(d.shape, np.int32)
    assert lib.tiab_dbg_watershed(d.ctypes.data, m.ctypes.data, k.ctypes.data, d.shape[0], d.shape[1], out.ctypes.data) == 0
    local = np.empty(d.shape, np.int32)
    assert lib.tiab_dbg_watershed_local(d.ctypes.data, m.ctypes.data, k.ctypes.data, d.shape[0], d.shape[1], local.ctypes.data) == 0
    assert np.array_equal(out, local)
    return out


def test_flood_on_pipeline_intermediates(built_lib):
    npm, hvm, _ = hovernet_head_maps(4, 164, seed=21)
    for i in range(4):
        for scale in (1.0, 0.5):
            dbg = {}
            want = po.proc_np_hv(npm[i], hvm[i], debug=dbg, scale_factor=scale)
            got = _flood(built_lib, dbg["dist"], dbg["marker"], dbg["blb"])
            assert np.array_equal(got, want), (i, scale, int((got != want).sum()))
            assert want.max() > 10


def test_flood_with_exact_ties(built_lib):
    rng = np.random.default_rng(3)
    for trial in range(300):
        h, w = int(rng.integers(4, 40)), int(rng.integers(4, 40))
        levels = int(rng.choice([1, 2, 3, 8]))                     # 1 level = one big plateau: every comparison is a tie on the value
        dist = rng.integers(0, levels, (h, w)).astype(np.float64) * 0.25 - 1.0
        mask = rng.random((h, w)) < rng.uniform(0.5, 1.0)
        markers = np.zeros((h, w), np.int32)
        for lab in range(1, int(rng.integers(2, 9))):               # multi-pixel markers (all age 0), some outside the mask
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            markers[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 4))] = lab
        want = po.watershed(dist, markers, mask)
        got = _flood(built_lib, dist, markers, mask)
        assert np.array_equal(got, want), (trial, h, w, levels)


def test_flood_known_answers(built_lib):
    img = np.array([[0.0, 1.0, 2.0, 5.0, 3.0, 1.5, 0.5]])
    markers = np.array([[1, 0, 0, 0, 0, 0, 2]])
    assert _flood(built_lib, img, markers, np.ones_like(markers)).tolist() == [[1, 1, 1, 1, 2, 2, 2]]
    img = np.zeros((1, 6))
    markers = np.array([[1, 0, 0, 0, 0, 2]])
    assert _flood(built_lib, img, markers, np.ones((1, 6))).tolist() == [[1, 1, 1, 2, 2, 2]]
    img = np.zeros((3, 5))
    markers = np.zeros((3, 5), int)
    markers[1, 0], markers[0, 4] = 7, 9
    mask = np.ones((3, 5), bool)
    mask[:, 2] = False
    mask[0, 4] = False
    out = _flood(built_lib, img, markers, mask)
    assert (out[:, :2] == 7).all() and (out[:, 2:] == 0).all()
