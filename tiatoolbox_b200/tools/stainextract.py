"""Stain-matrix extractors — same classes/semantics as tiatoolbox/tools/stainextract.py,
computed by libtiab200.so on the GPU (no CPU fallback).

Reference: tiatoolbox/tools/stainextract.py:13-68 (helpers), :71-101 CustomExtractor,
:104-137 RuifrokExtractor, :141-227 MacenkoExtractor, :230-322 VahadaneExtractor.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from tiatoolbox_b200 import _lib
from tiatoolbox_b200.utils.misc import luminosity_to_y_threshold

logger = logging.getLogger(__name__)

RUIFROK_MATRIX = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11]])


def vectors_in_correct_direction(e_vectors: np.ndarray) -> np.ndarray:
    """Point the eigen vectors in the right direction (stainextract.py:13-30)."""
    if e_vectors[0, 0] < 0:
        e_vectors[:, 0] *= -1
    if e_vectors[0, 1] < 0:
        e_vectors[:, 1] *= -1
    return e_vectors


def h_and_e_in_right_order(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    """H first (stainextract.py:33-50)."""
    if v1[0] > v2[0]:
        return np.array([v1, v2])
    return np.array([v2, v1])


def dl_output_for_h_and_e(dictionary: np.ndarray) -> np.ndarray:
    """Order dictionary-learning output as H, E (stainextract.py:53-68)."""
    if dictionary[0, 0] < dictionary[1, 0]:
        return dictionary[[1, 0], :]
    return dictionary


def _raise_for_status(status: int) -> None:
    if status == _lib.STAIN_EMPTY_TISSUE_MASK:
        msg = "Empty tissue mask computed."  # utils/misc.py:286-288
        raise ValueError(msg)
    if status == _lib.STAIN_DEGENERATE:
        msg = "Degenerate stain estimate (fewer than 2 tissue pixels or singular stain matrix)."
        raise ValueError(msg)
    if status == _lib.STAIN_UNUSED_ATOM:
        msg = "Vahadane dictionary atom unused (the reference re-samples a random atom here)."
        raise ValueError(msg)


class _ExtractorBase:
    """Shared plumbing: the C-ABI parameter block of this extractor."""

    method = "custom"
    #: CUDA device index the extraction runs on (a StainNormalizer sets it to its own `device` in fit / transform)
    device = 0

    def _params(self) -> _lib.StainParams:
        p = _lib.StainParams()
        p.method = _lib.METHOD_ID[self.method]
        p.y_threshold = luminosity_to_y_threshold(getattr(self, "_luminosity_threshold", 0.8))
        p.angular_percentile = float(getattr(self, "_angular_percentile", 99))
        p.regularizer = float(getattr(self, "_regularizer", 0.1))
        fixed = getattr(self, "_fixed_matrix", None)
        if fixed is not None:
            for i, v in enumerate(np.asarray(fixed, dtype=np.float64)[:2].reshape(-1)):
                p.fixed_matrix[i] = float(v)
        return p

    def _device_stain_matrix(self, img: np.ndarray, device: int | None = None) -> np.ndarray:
        device = self.device if device is None else device
        img = np.ascontiguousarray(np.asarray(img).astype("uint8"))  # "ensure input image is uint8"
        if img.ndim != 3 or img.shape[2] != 3:
            msg = "img must be HxWx3"
            raise ValueError(msg)
        h = _lib.handle(device)
        m = np.zeros(6, np.float64)
        mc = np.zeros(2, np.float64)
        st = C.c_int(0)
        p = self._params()
        rc = _lib.lib().tiab_stain_fit_host(
            h, C.byref(p), img.ctypes.data, img.shape[0], img.shape[1], _lib.dptr(m), _lib.dptr(mc), C.byref(st)
        )
        _lib.check(h, rc)
        _raise_for_status(st.value)
        return m.reshape(2, 3)


class CustomExtractor(_ExtractorBase):
    """Get the user-defined stain matrix (stainextract.py:71-101)."""

    method = "custom"

    def __init__(self, stain_matrix: np.ndarray) -> None:
        self.stain_matrix = stain_matrix
        if self.stain_matrix.shape not in [(2, 3), (3, 3)]:
            msg = "Stain matrix must have shape (2, 3) or (3, 3)."
            raise ValueError(msg)
        self._fixed_matrix = self.stain_matrix

    def get_stain_matrix(self, _: np.ndarray) -> np.ndarray:
        return self.stain_matrix


class RuifrokExtractor(_ExtractorBase):
    """Ruifrok & Johnston colour-deconvolution matrix (stainextract.py:104-137)."""

    method = "ruifrok"

    def __init__(self) -> None:
        self.__stain_matrix = RUIFROK_MATRIX.copy()
        self._fixed_matrix = self.__stain_matrix

    def get_stain_matrix(self, _: np.ndarray) -> np.ndarray:
        return self.__stain_matrix.copy()


class MacenkoExtractor(_ExtractorBase):
    """Macenko stain extractor (stainextract.py:141-227) on the GPU."""

    method = "macenko"

    def __init__(self, luminosity_threshold: float = 0.8, angular_percentile: float = 99) -> None:
        self._luminosity_threshold = luminosity_threshold
        self._angular_percentile = angular_percentile

    def get_stain_matrix(self, img: np.ndarray) -> np.ndarray:
        return self._device_stain_matrix(img)


class VahadaneExtractor(_ExtractorBase):
    """Vahadane stain extractor (stainextract.py:230-322) on the GPU."""

    method = "vahadane"

    def __init__(self, luminosity_threshold: float = 0.8, regularizer: float = 0.1) -> None:
        logger.warning(
            "Vahadane stain extraction/normalization algorithms are unstable "
            "after the update to `dictionary learning` algorithm in "
            "scikit-learn > v0.23.0 (see issue #382). Please be advised and "
            "consider using other stain extraction (normalization) algorithms.",
            stacklevel=2,
        )
        self._luminosity_threshold = luminosity_threshold
        self._regularizer = regularizer

    def get_stain_matrix(self, img: np.ndarray) -> np.ndarray:
        return self._device_stain_matrix(img)
