"""Seeded synthetic inputs for tests and bench.

H&E-like Beer-Lambert tiles so that the tissue mask, covariance and percentiles of the
stain path are non-degenerate.  Pure numpy; no reference code involved.
"""
from __future__ import annotations

import numpy as np

_HE = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11]])


def he_tiles(n: int, h: int = 256, w: int = 256, seed: int = 0, smooth: bool = False) -> np.ndarray:
    """(n, h, w, 3) uint8 H&E-like tiles.  ``seed`` fixes everything."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):  # tile-wise so that memory stays bounded for large n
        c = rng.gamma(2.0, 0.35, size=(h, w, 2)) * (rng.random((h, w, 1)) < 0.7)
        od = c @ _HE + rng.normal(0.0, 0.02, size=(h, w, 3))
        out[i] = np.clip(255.0 * np.exp(-od), 0, 255).astype(np.uint8)
    return out


def he_target(h: int = 512, w: int = 512, seed: int = 1234) -> np.ndarray:
    """A synthetic fit() target with a different stain appearance than he_tiles()."""
    rng = np.random.default_rng(seed)
    he = np.array([[0.55, 0.76, 0.35], [0.12, 0.85, 0.50]])
    c = rng.gamma(2.5, 0.4, size=(h, w, 2)) * (rng.random((h, w, 1)) < 0.8)
    od = c @ he + rng.normal(0.0, 0.02, size=(h, w, 3))
    return np.clip(255.0 * np.exp(-od), 0, 255).astype(np.uint8)


def hovernet_head_maps(n: int, size: int = 164, seed: int = 0):
    """Synthetic HoVer-Net head outputs: K ~ U{20..60} discs/ellipses per map
    (overlaps allowed), np = 0.95 inside / 0.05 outside + N(0, 0.02), hv = per-instance horizontal/vertical
    offsets from the instance centroid scaled to [-1, 1] + N(0, 0.01), tp = per-instance class U{1..5}.

    Returns (np_map (n,size,size,1), hv_map (n,size,size,2), tp_map (n,size,size,1)) float32."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    np_map = np.empty((n, size, size, 1), np.float32)
    hv_map = np.zeros((n, size, size, 2), np.float32)
    tp_map = np.zeros((n, size, size, 1), np.float32)
    for i in range(n):
        inst = np.zeros((size, size), np.int32)
        k = int(rng.integers(20, 61))
        for j in range(1, k + 1):
            cy, cx = rng.uniform(0, size, 2)
            ra, rb = rng.uniform(4, 12, 2)
            th = rng.uniform(0, np.pi)
            dx, dy = xx - cx, yy - cy
            u = (dx * np.cos(th) + dy * np.sin(th)) / ra
            v = (-dx * np.sin(th) + dy * np.cos(th)) / rb
            inst[(u * u + v * v) <= 1.0] = j
        fg = inst > 0
        np_map[i, ..., 0] = np.where(fg, 0.95, 0.05) + rng.normal(0, 0.02, (size, size))
        for j in np.unique(inst)[1:]:
            m = inst == j
            ys, xs = np.nonzero(m)
            cy, cx = ys.mean(), xs.mean()
            hx = xs - cx
            hy = ys - cy
            if hx.max() > 0 or hx.min() < 0:
                hx = np.where(hx < 0, hx / max(-hx.min(), 1e-6), hx / max(hx.max(), 1e-6))
            if hy.max() > 0 or hy.min() < 0:
                hy = np.where(hy < 0, hy / max(-hy.min(), 1e-6), hy / max(hy.max(), 1e-6))
            hv_map[i, ys, xs, 0] = hx
            hv_map[i, ys, xs, 1] = hy
            tp_map[i, ys, xs, 0] = float(rng.integers(1, 6))
        hv_map[i] += rng.normal(0, 0.01, (size, size, 2)).astype(np.float32)
    return np_map, hv_map, tp_map
