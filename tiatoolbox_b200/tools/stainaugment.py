"""Stain augmentation — drop-in for tiatoolbox/tools/stainaugment.py (`StainAugmentor`, :15-260) on a B200.

`fit` keeps the image's stain concentrations (float32, H*W x 2) and tissue mask in HBM; every `augment()` is ONE elementwise kernel
(csrc/augment.cu) over them, so generating many augmented instances of the same image — the reference's stand-alone use
(stainaugment.py:86-97) — never leaves the device until the uint8 result is copied back.

The reference class derives from albumentations' `ImageOnlyTransform`; when albumentations is importable this class does too (so it still
slots into `A.Compose`), otherwise it is a plain callable with the same `apply` / `get_params` / `__call__(image=...)` surface.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from tiatoolbox_b200 import _lib
from tiatoolbox_b200.tools.stainnorm import get_normalizer
from tiatoolbox_b200.utils.misc import get_luminosity_tissue_mask_device

try:  # pragma: no cover - albumentations is optional
    from albumentations.core.transforms_interface import ImageOnlyTransform as _Base
except ImportError:
    class _Base:  # the two constructor arguments albumentations transforms share
        def __init__(self, always_apply: bool = False, p: float = 0.5) -> None:  # noqa: FBT001, FBT002
            self.always_apply = always_apply
            self.p = p

        def __call__(self, *, image: np.ndarray, force_apply: bool = False, **_kw) -> dict:  # noqa: FBT001, FBT002
            if force_apply or self.always_apply or np.random.default_rng().random() < self.p:
                return {"image": self.apply(image)}
            return {"image": image}


class StainAugmentor(_Base):
    """Stain augmentation using a predefined stain matrix or a stain extraction method (stainaugment.py:15-141).

    Args: `method` ("vahadane" default, or "macenko"), `stain_matrix` (pre-extracted 2x3 matrix or None), `sigma1` / `sigma2` (ranges of the
    concentration scale alpha in [1 - sigma1, 1 + sigma1] and shift beta in [-sigma2, sigma2]), `p`, `augment_background`, `always_apply`.
    Attributes: `stain_normalizer`, `stain_matrix`, `source_concentrations` (device tensor; `.cpu().numpy()` gives the reference's array in
    float32), `n_stains`, `tissue_mask`, `alpha`, `beta`.
    """

    def __init__(self, method: str = "vahadane", stain_matrix: np.ndarray | None = None, sigma1: float = 0.4, sigma2: float = 0.2,
                 p: float = 0.5, *, augment_background: bool = False, always_apply: bool = False) -> None:
        super().__init__(always_apply=always_apply, p=p)
        self.augment_background = augment_background
        self.sigma1 = sigma1
        self.sigma2 = sigma2
        self.method = method
        self.stain_matrix = stain_matrix
        if self.method.lower() not in {"macenko", "vahadane"}:
            msg = f"Unsupported stain extractor method {self.method!r} for StainAugmentor. Choose either 'vahadane' or 'macenko'."
            raise ValueError(msg)
        self.stain_normalizer = get_normalizer(self.method.lower())
        self.alpha: float
        self.beta: float
        self.img_shape: tuple[int, ...]
        self.tissue_mask = None
        self.n_stains: int
        self.source_concentrations = None
        self.device = 0

    def fit(self, img: np.ndarray, threshold: float = 0.85) -> None:
        """Extract what `augment` needs (stainaugment.py:143-169): stain matrix (unless given), concentrations, tissue mask."""
        import torch

        img = np.ascontiguousarray(np.asarray(img).astype("uint8"))
        if img.ndim != 3 or img.shape[2] != 3:  # noqa: PLR2004
            msg = "img must be HxWx3"
            raise ValueError(msg)
        if self.stain_matrix is None:
            self.stain_normalizer.fit(img)
            self.stain_matrix = self.stain_normalizer.stain_matrix_target
        m = np.ascontiguousarray(np.asarray(self.stain_matrix, dtype=np.float64))
        if m.shape != (2, 3):
            msg = "StainAugmentor on the GPU supports 2x3 stain matrices."
            raise ValueError(msg)
        dev = torch.device("cuda", self.device)
        h = _lib.handle(self.device)
        x = torch.from_numpy(img).to(dev)
        conc = torch.empty((img.shape[0] * img.shape[1], 2), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(h, _lib.lib().tiab_stain_concentrations(h, x.data_ptr(), 1, img.shape[0], img.shape[1], _lib.dptr(m), conc.data_ptr(), stream))
        self.source_concentrations = conc
        self.n_stains = 2
        if not self.augment_background:
            self.tissue_mask = get_luminosity_tissue_mask_device(x[None], threshold=threshold)[0].reshape(-1)
        self.img_shape = img.shape

    def augment(self) -> np.ndarray:
        """One augmented instance of the fitted image (stainaugment.py:171-206)."""
        import torch

        if self.source_concentrations is None:
            msg = "fit() must be called before augment()."
            raise RuntimeError(msg)
        alphas, betas = [], []
        for _ in range(self.n_stains):
            self.get_params()
            alphas.append(float(self.alpha))
            betas.append(float(self.beta))
        conc = self.source_concentrations
        dev = conc.device
        out = torch.empty((conc.shape[0], 3), dtype=torch.uint8, device=dev)
        h = _lib.handle(dev.index or 0)
        m = np.ascontiguousarray(np.asarray(self.stain_matrix, dtype=np.float64))
        a, b = np.asarray(alphas, np.float64), np.asarray(betas, np.float64)
        mask_ptr = None if self.augment_background else self.tissue_mask.data_ptr()
        _lib.check(h, _lib.lib().tiab_stain_augment(h, conc.data_ptr(), mask_ptr, conc.shape[0], _lib.dptr(m), _lib.dptr(a), _lib.dptr(b),
                                                    out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return out.cpu().numpy().reshape(self.img_shape)

    def apply(self, img: np.ndarray, **params: dict) -> np.ndarray:  # noqa: ARG002
        """`fit(img, threshold=0.85)` then `augment()` (stainaugment.py:208-229)."""
        self.fit(img, threshold=0.85)
        return self.augment()

    def get_params(self) -> dict:
        """Draw alpha ~ U[1 - sigma1, 1 + sigma1] and beta ~ U[-sigma2, sigma2] (stainaugment.py:231-236)."""
        rng = np.random.default_rng()
        self.alpha = rng.uniform(1 - self.sigma1, 1 + self.sigma1)
        self.beta = rng.uniform(-self.sigma2, self.sigma2)
        return {}

    def get_params_dependent_on_targets(self, params: dict) -> dict:  # noqa: ARG002
        return {}

    @staticmethod
    def get_transform_init_args_names(**kwargs: dict) -> tuple[str, ...]:  # noqa: ARG004
        return "method", "stain_matrix", "sigma1", "sigma2", "augment_background"
