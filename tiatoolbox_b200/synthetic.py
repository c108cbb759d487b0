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
