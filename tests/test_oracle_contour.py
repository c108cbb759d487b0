"""The device's border-following code (csrc/contour_math.cuh, compiled for the host behind tiab_dbg_trace_contour) against
cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] — the call HoVerNet.get_instance_info makes per nucleus (hovernet.py:685-693)
— on hand-built shapes and a few thousand random 4-connected components (what a watershed instance is)."""
from __future__ import annotations

import cv2
import numpy as np
from scipy import ndimage


def _mine(lib, lab: np.ndarray, idv: int, cap: int | None = None) -> tuple[int, np.ndarray]:
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    cap = 2 * lab.size + 4 if cap is None else cap
    out = np.zeros(2 * max(cap, 1), np.int32)
    n = lib.tiab_dbg_trace_contour(lab.ctypes.data, lab.shape[0], lab.shape[1], int(idv), out.ctypes.data, cap)
    return n, out[: 2 * min(n, cap)].reshape(-1, 2)


def _ref(comp: np.ndarray) -> np.ndarray:
    ys, xs = np.nonzero(comp)
    y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    c, _ = cv2.findContours(comp[y0:y1, x0:x1].astype(np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    return c[0][:, 0, :] + np.array([x0, y0])


def test_known_shapes(built_lib):
    m = np.zeros((6, 7), np.int32)
    m[1:4, 1:5] = 5
    assert _mine(built_lib, m, 5)[1].tolist() == [[1, 1], [1, 3], [4, 3], [4, 1]]                 # rectangle: 4 corners, down first
    ring = np.zeros((7, 7), np.int32)
    ring[1:6, 1:6] = 2
    ring[3, 3] = 0
    assert _mine(built_lib, ring, 2)[1].tolist() == [[1, 1], [1, 5], [5, 5], [5, 1]]              # the hole is not part of [0][0]
    for pts, want in (([(2, 2)], [[2, 2]]), ([(2, 2), (2, 3)], [[2, 2], [3, 2]]), ([(1, 1), (1, 2), (2, 1)], [[1, 1], [1, 2], [2, 1]])):
        m = np.zeros((5, 5), np.int32)
        for y, x in pts:
            m[y, x] = 9
        assert _mine(built_lib, m, 9)[1].tolist() == want
    assert _mine(built_lib, np.zeros((4, 4), np.int32), 1)[0] == 0                                 # label absent
    n, pts = _mine(built_lib, ring, 2, cap=2)                                                      # counting works past the capacity
    assert n == 4 and pts.tolist() == [[1, 1], [1, 5]]
    edge = np.full((3, 4), 4, np.int32)                                                            # touches every image border
    assert _mine(built_lib, edge, 4)[1].tolist() == [[0, 0], [0, 2], [3, 2], [3, 0]]


def test_fuzz_against_cv2(built_lib):
    rng = np.random.default_rng(0)
    checked = 0
    for trial in range(3000):
        h, w = int(rng.integers(3, 24)), int(rng.integers(3, 24))
        kind = trial % 4
        if kind == 0:
            m = rng.random((h, w)) < rng.uniform(0.3, 0.9)
        elif kind == 1:
            m = np.zeros((h, w), bool)
            yy, xx = np.mgrid[:h, :w]
            for _ in range(int(rng.integers(1, 4))):
                cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(1, 7)
                m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        elif kind == 2:   # thin 4-connected random walks: borders traversed in both directions
            m = np.zeros((h, w), bool)
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            for _ in range(int(rng.integers(1, 60))):
                m[y, x] = True
                d = int(rng.integers(0, 4))
                y = min(max(y + [0, 0, 1, -1][d], 0), h - 1)
                x = min(max(x + [1, -1, 0, 0][d], 0), w - 1)
        else:
            m = rng.random((h, w)) < 0.75
        lab4, n = ndimage.label(m)
        if n == 0:
            continue
        comp = lab4 == int(rng.integers(1, n + 1))
        lab = np.where(comp, 7, np.where(m, 3, 0)).astype(np.int32)     # other labels around it, like neighbouring nuclei
        count, got = _mine(built_lib, lab, 7)
        want = _ref(comp)
        assert count == len(want) and np.array_equal(got, want), (trial, comp.astype(int))
        checked += 1
    assert checked > 2500
