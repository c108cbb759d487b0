"""Host-side helpers of the stain path (mirrors of tiatoolbox/utils/misc.py pieces)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from tiatoolbox_b200.utils.exceptions import FileNotSupportedError


def load_stain_matrix(stain_matrix_input):
    """Load a stain matrix as a numpy array (tiatoolbox/utils/misc.py:218-258)."""
    if isinstance(stain_matrix_input, (str, Path)):
        suffixes = Path(stain_matrix_input).suffixes
        if not suffixes or suffixes[-1] not in [".csv", ".npy"]:
            msg = "If supplying a path to a stain matrix, use either a npy or a csv file"
            raise FileNotSupportedError(msg)
        if suffixes[-1] == ".csv":
            import pandas as pd

            return pd.read_csv(stain_matrix_input).to_numpy()
        return np.load(str(stain_matrix_input))
    if isinstance(stain_matrix_input, np.ndarray):
        return stain_matrix_input
    msg = "Stain_matrix must be either a path to npy/csv file or a numpy array"
    raise TypeError(msg)


def luminosity_to_y_threshold(threshold: float) -> int:
    """Map get_luminosity_tissue_mask's ``threshold`` (misc.py:282-283, ``L/255.0 < threshold``
    on cv2's 8-bit Lab L) to the fixed-point luminance bound the kernels test: tissue <=>
    Y <= returned value, where L = clip((296*tabc[Y] - 1336934 + 16384) >> 15, 0, 255) is
    OpenCV's fixed-point L for Y in 0..2040 (monotone).  -1 if no L qualifies."""
    t = np.arange(2041, dtype=np.float64) / 2040.0
    f = np.where(t < 216.0 / 24389.0, t * 841.0 / 108.0 + 16.0 / 116.0, np.cbrt(t))
    tabc = np.rint(32768.0 * f).astype(np.int64)
    l_of_y = np.clip((296 * tabc - 1336934 + 16384) >> 15, 0, 255)
    ok = (l_of_y / 255.0) < threshold
    return int(np.nonzero(ok)[0].max()) if ok.any() else -1
