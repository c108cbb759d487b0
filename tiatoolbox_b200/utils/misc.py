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


def get_luminosity_tissue_mask_device(imgs, threshold: float = 0.8):
    """Tissue masks of an (N, H, W, 3) uint8 CUDA tensor on the device: (N, H, W) uint8 tensor (1 = tissue).  Raises ValueError
    "Empty tissue mask computed." like utils/misc.py:286-288 when an image has no tissue pixel."""
    import ctypes as C

    import torch

    from tiatoolbox_b200 import _lib

    x = imgs.contiguous()
    if not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:  # noqa: PLR2004
        msg = "expected a (N,H,W,3) uint8 CUDA tensor"
        raise ValueError(msg)
    n, hh, ww, _ = x.shape
    dev = x.device
    mask = torch.empty((n, hh, ww), dtype=torch.uint8, device=dev)
    status = torch.zeros((n,), dtype=torch.int32, device=dev)
    p = _lib.StainParams()
    p.method = _lib.METHOD_ID["macenko"]
    p.y_threshold = luminosity_to_y_threshold(threshold)
    p.angular_percentile, p.regularizer = 99.0, 0.1
    h = _lib.handle(dev.index or 0)
    _lib.check(h, _lib.lib().tiab_stain_tissue_mask(h, C.byref(p), x.data_ptr(), mask.data_ptr(), n, hh, ww, status.data_ptr(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
    if bool((status == _lib.STAIN_EMPTY_TISSUE_MASK).any().item()):
        msg = "Empty tissue mask computed."
        raise ValueError(msg)
    return mask


def get_luminosity_tissue_mask(img: np.ndarray, threshold: float) -> np.ndarray:
    """utils/misc.py:261-290: boolean tissue mask of an RGB uint8 image (contrast enhancement -> cv2 8-bit Lab L / 255 < threshold), on the device."""
    import torch

    arr = np.ascontiguousarray(np.asarray(img).astype("uint8"))
    if arr.ndim != 3 or arr.shape[2] != 3:  # noqa: PLR2004
        msg = "img must be HxWx3"
        raise ValueError(msg)
    return get_luminosity_tissue_mask_device(torch.from_numpy(arr).cuda()[None], threshold=threshold)[0].cpu().numpy().astype(bool)
