"""ModelABC — the model contract the engines rely on (mirrors tiatoolbox/models/models_abc.py:87-256).

A B200 model is a weight container (flat state dict with the reference's keys) + an fp32 PyTorch
restatement of the forward (`forward`, used as the parity reference and for `device="cpu"` debugging of
*weights*, never as a silent fallback of the hot path) + device plans built lazily per (device, batch,
precision).  `infer_batch(model, batch_data, *, device)` is static like the reference's and is what the
engines fetch (engine_abc.py:389-395, 555).
"""
from __future__ import annotations

from collections import OrderedDict
from pathlib import Path
from typing import Callable

import numpy as np
import torch


class ModelABC(torch.nn.Module):
    """Base class: pre/post-processing hooks, weights, plan cache."""

    #: "fp16x3" (split-fp16, meets 1e-3 end to end) or "fp16" (single pass, TF32-equivalent operands)
    precision = "fp16x3"

    def __init__(self) -> None:
        super().__init__()
        self._postproc = self.postproc
        self._preproc = self.preproc
        self.class_dict = None
        self._sd: OrderedDict = OrderedDict()
        self._plans: dict = {}
        self._device = torch.device("cuda", 0)

    # ---- hooks (models_abc.py:133-201) ----
    @staticmethod
    def preproc(image: np.ndarray) -> np.ndarray:
        return image

    @staticmethod
    def postproc(image: np.ndarray) -> np.ndarray:
        return image

    @property
    def preproc_func(self) -> Callable:
        return self._preproc

    @preproc_func.setter
    def preproc_func(self, func: Callable | None) -> None:
        if func is not None and not callable(func):
            msg = f"{func} is not callable!"
            raise ValueError(msg)
        self._preproc = self.preproc if func is None else func

    @property
    def postproc_func(self) -> Callable:
        return self._postproc

    @postproc_func.setter
    def postproc_func(self, func: Callable | None) -> None:
        if func is not None and not callable(func):
            msg = f"{func} is not callable!"
            raise ValueError(msg)
        self._postproc = self.postproc if func is None else func

    # ---- weights ----
    def state_dict(self, *args, **kwargs) -> OrderedDict:  # noqa: ARG002
        return OrderedDict((k, v) for k, v in self._sd.items())

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):  # noqa: ARG002
        missing = [k for k in self._sd if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._sd]
        if strict and (missing or unexpected):
            msg = f"Error(s) in loading state_dict: missing keys {missing[:5]}..., unexpected keys {unexpected[:5]}..."
            raise RuntimeError(msg)
        for k, v in state_dict.items():
            if k in self._sd:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    msg = f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(self._sd[k].shape)}"
                    raise RuntimeError(msg)
                self._sd[k] = v.detach().clone().cpu()
        self._plans.clear()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_weights_from_file(self, weights: str | Path):
        """models_abc.py:239-256 / load_torch_model :26-44."""
        pretrained = torch.load(weights, map_location="cpu")
        self.load_state_dict(pretrained, strict=True)
        return self

    def to(self, device="cuda", *args, **kwargs):  # noqa: ARG002
        dev = torch.device(device)
        if dev.type != "cuda":
            msg = "tiatoolbox_b200 models run on CUDA devices only (there is no CPU fallback of the hot path)"
            raise ValueError(msg)
        self._device = torch.device("cuda", dev.index or 0)
        return self

    def eval(self):
        return self

    # ---- plans ----
    def _build_plan(self, batch: int, device: torch.device, shape):  # pragma: no cover - abstract
        raise NotImplementedError

    def plan_for(self, batch: int, device, shape=None):
        dev = torch.device(device)
        dev = torch.device("cuda", dev.index or 0)
        key = (batch, str(dev), self.precision, tuple(shape) if shape is not None else None)
        if key not in self._plans:
            plan = self._build_plan(batch, dev, shape)
            plan.capture()
            self._plans[key] = plan
        return self._plans[key]

    @staticmethod
    def _nchw_to_u8_nhwc(x: torch.Tensor, device) -> torch.Tensor:
        """`forward()` input (NCHW, pixel values 0..255 as the reference's forward expects) -> the plans' uint8 NHWC batch.  Non-integer
        values cannot come out of the reference's uint8 data path and are rejected rather than silently rounded."""
        x = torch.as_tensor(x)
        if x.dim() != 4:  # noqa: PLR2004
            msg = f"expected an NCHW batch, got shape {tuple(x.shape)}"
            raise ValueError(msg)
        if x.dtype != torch.uint8:
            r = x.round()
            if float((x - r).abs().max()) > 1e-3 or float(r.min()) < 0 or float(r.max()) > 255:  # noqa: PLR2004
                msg = "the B200 forward takes uint8 pixel values (0..255): there is no PyTorch-eager forward for arbitrary float inputs"
                raise ValueError(msg)
            x = r.to(torch.uint8)
        return x.permute(0, 2, 3, 1).contiguous().to(device)

    @staticmethod
    def _as_device_u8(batch_data, device) -> torch.Tensor:
        """uint8 NHWC batch on the device (float inputs in [0, 255] are rounded like the reference's
        DataLoader output would be cast by the kernels)."""
        x = batch_data if isinstance(batch_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(batch_data))
        if x.dtype != torch.uint8:
            x = x.round().clamp(0, 255).to(torch.uint8)
        return x.to(device, non_blocking=True).contiguous()
