"""Stain normalisation — drop-in for tiatoolbox/tools/stainnorm.py on a B200.

Same factory, classes, attributes, argument meaning and exceptions as the reference
(tiatoolbox/tools/stainnorm.py:19-219, 370-425); the arithmetic runs in libtiab200.so
(fused sm_100a kernel, see csrc/stain.cu).  New: ``transform_batch`` for (N,H,W,3) input
(numpy or CUDA torch tensors), which is what the patch engines call.

Conscious difference: the reference's ``rgb2od`` overwrites zeros with ones in the
*caller's* array (utils/transforms.py:229-230); this implementation never mutates its input.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from tiatoolbox_b200 import _lib
from tiatoolbox_b200.tools.stainextract import (
    CustomExtractor,
    MacenkoExtractor,
    RuifrokExtractor,
    VahadaneExtractor,
    _raise_for_status,
)
from tiatoolbox_b200.utils.exceptions import MethodNotSupportedError
from tiatoolbox_b200.utils.misc import load_stain_matrix
from tiatoolbox_b200.utils.transforms import od2rgb


def _current_device() -> int:
    try:
        import torch

        return int(torch.cuda.current_device()) if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


def _concentrations_three_stains(img_u8, stain_matrix: np.ndarray):
    """`lstsq(M.T, od.T).T` of stainnorm.py:64-66 for a 3x3 `M` on the tensor's device, float64: the minimum-norm least-squares solution is
    `od @ pinv(M)`; `od` as `rgb2od` (transforms.py:229-231: zeros -> 1, floor 1e-6)."""
    import torch

    pinv = torch.from_numpy(np.linalg.pinv(np.asarray(stain_matrix, np.float64))).to(img_u8.device)
    x = img_u8.reshape(-1, 3).to(torch.float64).clamp_(min=1.0)
    od = (-torch.log(x / 255.0)).clamp_(min=1e-6)
    return od @ pinv


class StainNormalizer:
    """Stain normalization base class (stainnorm.py:19-113).

    Attributes:
        extractor: method specific stain extractor.
        stain_matrix_target (ndarray): 2x3 stain matrix of the target.
        target_concentrations (ndarray): stain concentrations of the target (computed lazily).
        maxC_target (ndarray): 1x2, 99th percentile of each stain.
        stain_matrix_target_RGB (ndarray): target stain matrix in RGB.
    """

    def __init__(self) -> None:
        self.extractor: CustomExtractor | MacenkoExtractor | RuifrokExtractor | VahadaneExtractor
        self.stain_matrix_target: np.ndarray
        self.maxC_target = None
        self.stain_matrix_target_RGB: np.ndarray
        self._target = None
        self._target_conc = None
        self.device = _current_device()      # one process per GPU: the thread's current CUDA device (torch.cuda.set_device(local_rank))

    # -- reference API ------------------------------------------------------------------------
    @staticmethod
    def get_concentrations(img: np.ndarray, stain_matrix: np.ndarray, device: int | None = None) -> np.ndarray:
        """Estimate the (H*W, 2) concentration matrix (stainnorm.py:49-66) on the GPU (`device`: default = the current CUDA device)."""
        import torch

        device = _current_device() if device is None else int(device)

        img = np.ascontiguousarray(np.asarray(img, dtype=np.uint8))
        m = np.ascontiguousarray(np.asarray(stain_matrix, dtype=np.float64))
        dev = torch.device("cuda", device)
        if m.shape == (3, 3):
            # three-stain user matrix (CustomExtractor accepts it, stainextract.py:73-77; `fit` cannot use it, but this static method can):
            # not on any engine's path, so it is the same least-squares solve spelled with device tensor ops instead of a kernel
            return _concentrations_three_stains(torch.from_numpy(img).to(dev), m).cpu().numpy()
        if m.shape != (2, 3):
            msg = "Stain matrix must have shape (2, 3) or (3, 3)."
            raise ValueError(msg)
        h = _lib.handle(device)
        t_in = torch.from_numpy(img).to(dev)
        t_out = torch.empty((img.shape[0] * img.shape[1], 2), dtype=torch.float32, device=dev)
        rc = _lib.lib().tiab_stain_concentrations(
            h, t_in.data_ptr(), 1, img.shape[0], img.shape[1], _lib.dptr(m), t_out.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream,
        )
        _lib.check(h, rc)
        return t_out.cpu().numpy().astype(np.float64)

    @property
    def target_concentrations(self) -> np.ndarray:
        if self._target_conc is None and self._target is not None:
            self._target_conc = self.get_concentrations(self._target, self.stain_matrix_target, self.device)
        return self._target_conc

    def fit(self, target: np.ndarray) -> None:
        """Fit to a target image (stainnorm.py:68-87)."""
        target = np.ascontiguousarray(np.asarray(target).astype("uint8"))
        if target.ndim != 3 or target.shape[2] != 3:
            msg = "target must be HxWx3"
            raise ValueError(msg)
        fixed = getattr(self.extractor, "stain_matrix", None)     # only a user matrix can have three rows (CustomExtractor)
        if fixed is not None and np.asarray(fixed).shape != (2, 3):
            # the reference fails in np.percentile(...).reshape((1, 2)) for 3-row matrices (:81-85)
            msg = "cannot reshape array of size 3 into shape (1,2)"
            raise ValueError(msg)
        self.extractor.device = self.device                        # extraction runs on this normaliser's GPU
        h = _lib.handle(self.device)
        m = np.zeros(6, np.float64)
        mc = np.zeros(2, np.float64)
        st = C.c_int(0)
        p = self.extractor._params()
        rc = _lib.lib().tiab_stain_fit_host(
            h, C.byref(p), target.ctypes.data, target.shape[0], target.shape[1], _lib.dptr(m), _lib.dptr(mc), C.byref(st)
        )
        _lib.check(h, rc)
        _raise_for_status(st.value)
        self.stain_matrix_target = m.reshape(2, 3)
        self.maxC_target = mc.reshape((1, 2))
        self._target = target
        self._target_conc = None
        self.stain_matrix_target_RGB = od2rgb(self.stain_matrix_target)

    def transform(self, img: np.ndarray) -> np.ndarray:
        """Transform one RGB uint8 image (stainnorm.py:89-113)."""
        img = np.asarray(img)
        out = self.transform_batch(img.astype("uint8")[None])
        return out[0].reshape(img.shape)

    # -- new: batched entry point ---------------------------------------------------------------
    def transform_batch(self, imgs, *, return_info: bool = False, raise_on_error: bool = True):
        """Normalise a batch (N,H,W,3) uint8.  numpy in -> numpy out (host round trip inside
        the C call); CUDA torch tensor in -> CUDA torch tensor out (no host copy)."""
        if self.maxC_target is None:
            msg = "fit() must be called before transform()."
            raise RuntimeError(msg)
        mt = np.ascontiguousarray(self.stain_matrix_target, dtype=np.float64).reshape(-1)
        mc = np.ascontiguousarray(self.maxC_target, dtype=np.float64).reshape(-1)
        p = self.extractor._params()
        L = _lib.lib()
        if _is_torch_cuda(imgs):
            import torch

            x = imgs.contiguous()
            if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
                msg = "expected a (N,H,W,3) uint8 tensor"
                raise ValueError(msg)
            dev = x.device
            h = _lib.handle(dev.index or 0)
            n, hh, ww, _ = x.shape
            out = torch.empty_like(x)
            msrc = torch.empty((n, 6), dtype=torch.float64, device=dev)
            maxc = torch.empty((n, 2), dtype=torch.float64, device=dev)
            status = torch.empty((n,), dtype=torch.int32, device=dev)
            rc = L.tiab_stain_transform_batch(
                h, C.byref(p), x.data_ptr(), out.data_ptr(), n, hh, ww, _lib.dptr(mt), _lib.dptr(mc),
                msrc.data_ptr(), maxc.data_ptr(), status.data_ptr(), None, torch.cuda.current_stream(dev).cuda_stream,
            )
            _lib.check(h, rc)
            if raise_on_error:
                bad = status.cpu().numpy()
                if bad.any():
                    _raise_for_status(int(bad[bad != 0][0]))
            if return_info:
                return out, {"stain_matrix_source": msrc.view(n, 2, 3), "maxC_source": maxc, "status": status}
            return out
        x = np.ascontiguousarray(np.asarray(imgs))
        if x.dtype != np.uint8 or x.ndim != 4 or x.shape[-1] != 3:
            msg = "expected a (N,H,W,3) uint8 array"
            raise ValueError(msg)
        n, hh, ww, _ = x.shape
        h = _lib.handle(self.device)
        out = np.empty_like(x)
        msrc = np.zeros((n, 6), np.float64)
        maxc = np.zeros((n, 2), np.float64)
        status = np.zeros((n,), np.int32)
        rc = L.tiab_stain_transform_host(
            h, C.byref(p), x.ctypes.data, out.ctypes.data, n, hh, ww, _lib.dptr(mt), _lib.dptr(mc),
            msrc.ctypes.data, maxc.ctypes.data, status.ctypes.data,
        )
        _lib.check(h, rc)
        if raise_on_error and status.any():
            _raise_for_status(int(status[status != 0][0]))
        if return_info:
            return out, {"stain_matrix_source": msrc.reshape(n, 2, 3), "maxC_source": maxc, "status": status}
        return out


class CustomNormalizer(StainNormalizer):
    """Stain normalisation using a user-defined stain matrix (stainnorm.py:116-140)."""

    def __init__(self, stain_matrix: np.ndarray) -> None:
        super().__init__()
        self.extractor = CustomExtractor(stain_matrix)


class RuifrokNormalizer(StainNormalizer):
    """Ruifrok & Johnston stain normalizer (stainnorm.py:143-168)."""

    def __init__(self) -> None:
        super().__init__()
        self.extractor = RuifrokExtractor()


class MacenkoNormalizer(StainNormalizer):
    """Macenko stain normalizer (stainnorm.py:171-194)."""

    def __init__(self) -> None:
        super().__init__()
        self.extractor = MacenkoExtractor()


class VahadaneNormalizer(StainNormalizer):
    """Vahadane stain normalizer (stainnorm.py:197-219)."""

    def __init__(self) -> None:
        super().__init__()
        self.extractor = VahadaneExtractor()


class ReinhardNormalizer(StainNormalizer):
    """Reinhard colour normalizer (stainnorm.py:222-367): match the mean / standard deviation of the 8-bit Lab channels to the target's.

    Attributes:
        target_means (tuple of float): mean of each Lab channel (L in [0, 100], a and b centred on 0) of the target.
        target_stds (tuple of float): standard deviation of each Lab channel of the target.

    The reference leans on OpenCV for both colour conversions and for the statistics; here the whole transform is integer device code
    (csrc/reinhard.cu, csrc/lab_math.cuh) that equals cv2's 8-bit conversions on all 2^24 inputs, so the output is the reference's
    uint8 image bit for bit unless a float32 rounding of the per-level map falls differently (never observed; the tests allow 1 LSB).
    Like the reference, an image with a constant Lab channel fails with ZeroDivisionError (stainnorm.py:277-283 divide by its std).
    """

    def __init__(self) -> None:
        super().__init__()
        self.target_means: tuple[float, float, float]
        self.target_stds: tuple[float, float, float]
        self._fitted = False

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _convert(img, inverse: int):
        """cv2.cvtColor(img, COLOR_RGB2LAB / COLOR_LAB2RGB) for uint8 (..., 3) data on the device."""
        import torch

        if _is_torch_cuda(img):
            x = img.contiguous()
            if x.dtype != torch.uint8 or x.shape[-1] != 3:
                msg = "expected a (..., 3) uint8 tensor"
                raise ValueError(msg)
            dev = x.device
            out = torch.empty_like(x)
            h = _lib.handle(dev.index or 0)
            if x.numel():
                _lib.check(h, _lib.lib().tiab_lab_convert(h, x.data_ptr(), out.data_ptr(), x.numel() // 3, inverse,
                                                          torch.cuda.current_stream(dev).cuda_stream))
            return out
        arr = np.ascontiguousarray(np.asarray(img).astype("uint8"))
        if arr.shape[-1] != 3:
            msg = "expected a (..., 3) uint8 array"
            raise ValueError(msg)
        return ReinhardNormalizer._convert(torch.from_numpy(arr).cuda(), inverse).cpu().numpy()

    @staticmethod
    def rgb2lab(img):
        """8-bit Lab image of an RGB uint8 image (numpy -> numpy, CUDA tensor -> CUDA tensor); equals cv2's COLOR_RGB2LAB."""
        return ReinhardNormalizer._convert(img, 0)

    @staticmethod
    def lab2rgb(img):
        """RGB uint8 image of an 8-bit Lab image; equals cv2's COLOR_LAB2RGB."""
        return ReinhardNormalizer._convert(img, 1)

    @staticmethod
    def lab_split(img: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """RGB uint8 -> float32 L in [0, 100], a and b in [-128, 127] (stainnorm.py:286-312)."""
        lab = ReinhardNormalizer.rgb2lab(np.asarray(img).astype("uint8")).astype(np.float32)
        chan1, chan2, chan3 = (np.ascontiguousarray(lab[..., i]) for i in range(3))
        chan1 /= np.asarray(2.55)
        chan2 -= np.asarray(128.0)
        chan3 -= np.asarray(128.0)
        return chan1, chan2, chan3

    @staticmethod
    def merge_back(chan1: np.ndarray, chan2: np.ndarray, chan3: np.ndarray) -> np.ndarray:
        """Separate Lab channels -> RGB uint8 (stainnorm.py:314-335; like the reference the channels are modified in place)."""
        chan1 *= 2.55
        chan2 += 128.0
        chan3 += 128.0
        lab = np.clip(np.stack((chan1, chan2, chan3), axis=-1), 0, 255).astype(np.uint8)
        return ReinhardNormalizer.lab2rgb(lab)

    def get_mean_std(self, img: np.ndarray):
        """Mean and standard deviation of each Lab channel (stainnorm.py:337-367) -> ((mL, ma, mb), (sL, sa, sb))."""
        import torch

        img = np.ascontiguousarray(np.asarray(img).astype("uint8"))
        if img.ndim != 3 or img.shape[2] != 3:  # noqa: PLR2004
            msg = "img must be HxWx3"
            raise ValueError(msg)
        dev = torch.device("cuda", self.device)
        x = torch.from_numpy(img).to(dev)
        h = _lib.handle(self.device)
        means, stds = np.zeros(3, np.float64), np.zeros(3, np.float64)
        _lib.check(h, _lib.lib().tiab_reinhard_fit(h, x.data_ptr(), img.shape[0], img.shape[1], _lib.dptr(means), _lib.dptr(stds),
                                                   torch.cuda.current_stream(dev).cuda_stream))
        return tuple(float(v) for v in means), tuple(float(v) for v in stds)

    # -- reference API ---------------------------------------------------------------------------
    def fit(self, target: np.ndarray) -> None:
        """Fit to a target image (stainnorm.py:251-260)."""
        self.target_means, self.target_stds = self.get_mean_std(target)
        self._fitted = True

    def transform(self, img: np.ndarray) -> np.ndarray:
        """Transform one RGB uint8 image (stainnorm.py:262-284)."""
        img = np.asarray(img)
        return self.transform_batch(img.astype("uint8")[None])[0].reshape(img.shape)

    def transform_batch(self, imgs, *, return_info: bool = False, raise_on_error: bool = True):
        """(N, H, W, 3) uint8 numpy array -> numpy array, or CUDA tensor -> CUDA tensor (no host copy)."""
        import torch

        if not self._fitted:
            msg = "fit() must be called before transform()."
            raise RuntimeError(msg)
        was_numpy = not _is_torch_cuda(imgs)
        if was_numpy:
            arr = np.ascontiguousarray(np.asarray(imgs))
            if arr.dtype != np.uint8 or arr.ndim != 4 or arr.shape[-1] != 3:  # noqa: PLR2004
                msg = "expected a (N,H,W,3) uint8 array"
                raise ValueError(msg)
            x = torch.from_numpy(arr).to(torch.device("cuda", self.device))
        else:
            x = imgs.contiguous()
            if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:  # noqa: PLR2004
                msg = "expected a (N,H,W,3) uint8 tensor"
                raise ValueError(msg)
        dev = x.device
        n, hh, ww, _ = x.shape
        out = torch.empty_like(x)
        stats = torch.empty((n, 6), dtype=torch.float64, device=dev)
        status = torch.empty((n,), dtype=torch.int32, device=dev)
        if n:
            h = _lib.handle(dev.index or 0)
            tm = np.ascontiguousarray(self.target_means, dtype=np.float64)
            ts = np.ascontiguousarray(self.target_stds, dtype=np.float64)
            _lib.check(h, _lib.lib().tiab_reinhard_transform_batch(h, x.data_ptr(), out.data_ptr(), n, hh, ww, _lib.dptr(tm), _lib.dptr(ts),
                                                                   stats.data_ptr(), status.data_ptr(),
                                                                   torch.cuda.current_stream(dev).cuda_stream))
            if raise_on_error and bool(status.any().item()):
                msg = "float division by zero"   # a Lab channel of the image is constant: stainnorm.py:277-283 divide by its std
                raise ZeroDivisionError(msg)
        res = out.cpu().numpy() if was_numpy else out
        if return_info:
            return res, {"source_means": stats[:, :3], "source_stds": stats[:, 3:], "status": status}
        return res


def set_reinhard_path(path: int, device: int = 0) -> None:
    """Kernel path of `ReinhardNormalizer.transform_batch` on this thread's handle of `device`: 0 = histogram + finalise + apply launches
    (default); 1 = experimental fused kernel (one CTA per image, the image resident in shared memory; images up to 256 x 256 x 3)."""
    h = _lib.handle(device)
    _lib.check(h, _lib.lib().tiab_reinhard_set_path(h, int(path)))


def get_normalizer(method_name: str, stain_matrix=None) -> StainNormalizer:
    """Return a :class:`StainNormalizer` by name (stainnorm.py:370-425)."""
    if method_name.lower() not in ["reinhard", "ruifrok", "macenko", "vahadane", "custom"]:
        raise MethodNotSupportedError

    if stain_matrix is not None and method_name.lower() != "custom":
        msg = '`stain_matrix` is only defined when using `method_name`="custom".'
        raise ValueError(msg)

    if method_name.lower() == "reinhard":
        return ReinhardNormalizer()
    if method_name.lower() == "ruifrok":
        return RuifrokNormalizer()
    if method_name.lower() == "macenko":
        return MacenkoNormalizer()
    if method_name.lower() == "vahadane":
        return VahadaneNormalizer()

    if stain_matrix is None:
        msg = '`stain_matrix` is None when using `method_name`="custom".'
        raise ValueError(msg)
    return CustomNormalizer(load_stain_matrix(stain_matrix))
