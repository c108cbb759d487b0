"""od2rgb / rgb2od for small arrays (stain-matrix visualisation), mirroring
tiatoolbox/utils/transforms.py:209-256.  Image-sized OD conversion lives in the kernels."""
from __future__ import annotations

import numpy as np


def rgb2od(img: np.ndarray) -> np.ndarray:
    img = np.where(img == 0, 1, img)
    return np.maximum(-1 * np.log(img / 255), 1e-6)


def od2rgb(od: np.ndarray) -> np.ndarray:
    od = np.maximum(od, 1e-6)
    return (255 * np.exp(-1 * od)).astype(np.uint8)
