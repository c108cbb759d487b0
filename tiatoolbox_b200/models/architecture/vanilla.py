"""CNNModel (torchvision backbone + linear classifier) for the B200 — `resnet18-kather100k` and friends.

Mirrors tiatoolbox/models/architecture/vanilla.py:256-359: same constructor, state_dict keys
(`feat_extract.<i>...`, `classifier.*`), forward semantics (trunk -> GAP -> Linear -> softmax) and
`infer_batch` / `postproc` contract.  Only the ResNet family is on the B200 path."""
from __future__ import annotations

import torch

from tiatoolbox_b200.models.architecture._params import ParamSpec
from tiatoolbox_b200.models.architecture.resnet_common import BASIC, BOTTLENECK, resnet_spec

# nn.Sequential(*list(model.children())[:-2]) -> indices of the torchvision ResNet children
_LAYERS = ("4", "5", "6", "7")
_STEM = ("0", "1")


def cnn_spec(backbone: str, num_classes: int) -> ParamSpec:
    if backbone not in BASIC and backbone not in BOTTLENECK:
        msg = f"Backbone `{backbone}` is not supported."
        raise ValueError(msg)
    sp = ParamSpec()
    c = resnet_spec(sp, "feat_extract.", backbone, 3, _LAYERS, _STEM)
    sp.linear("classifier", num_classes, c)
    return sp


# =================================================================================================
# Model class (API mirror of tiatoolbox/models/architecture/vanilla.py:256-359)
# =================================================================================================
import numpy as np  # noqa: E402

from tiatoolbox_b200.models.architecture._params import seeded_state_dict  # noqa: E402
from tiatoolbox_b200.models.models_abc import ModelABC  # noqa: E402


class CNNModel(ModelABC):
    """Backbone + linear classifier; softmax probabilities out."""

    #: scale applied to uint8 pixels before the network (the kather100k preproc is ToTensor(): x / 255)
    u8_scale = 1.0 / 255.0

    def __init__(self, backbone: str, num_classes: int = 1) -> None:
        super().__init__()
        self.backbone = backbone
        self.num_classes = num_classes
        self._sd = seeded_state_dict(cnn_spec(backbone, num_classes), seed=0)

    def forward(self, imgs: torch.Tensor) -> torch.Tensor:
        """vanilla.py:300-316 on the device plan: NCHW input as the dataset feeds it (ToTensor: uint8 / 255) -> softmax probabilities."""
        x = ModelABC._nchw_to_u8_nhwc(imgs / self.u8_scale, self._device)
        plan = self.plan_for(x.shape[0], self._device, x.shape[1:3])
        return plan.forward(x).clone()

    def _build_plan(self, batch, device, shape):
        from tiatoolbox_b200.models.resnet_plan import CNNPlan

        return CNNPlan(self._sd, backbone=self.backbone, num_classes=self.num_classes, batch=batch, device=device,
                       in_size=int(shape[0]), u8_scale=self.u8_scale, precision=self.precision)

    @staticmethod
    def postproc(image: np.ndarray) -> np.ndarray:
        """argmax along the last axis (vanilla.py:318-333)."""
        return np.argmax(image, axis=-1)

    @staticmethod
    def infer_batch_device(model: "CNNModel", batch_data, *, device) -> torch.Tensor:
        x = batch_data
        if isinstance(x, np.ndarray) and x.dtype != np.uint8 or isinstance(x, torch.Tensor) and x.dtype != torch.uint8:
            # the reference's dataset hands over ToTensor()-scaled floats in [0, 1]: undo the scaling exactly
            x = (torch.as_tensor(x) / model.u8_scale)
        x = ModelABC._as_device_u8(x, device)
        n = x.shape[0]
        plan_b = getattr(model, "_plan_batch", None) or n
        plan = model.plan_for(plan_b, device, x.shape[1:3])
        if n < plan_b:
            x = torch.cat([x, torch.zeros((plan_b - n, *x.shape[1:]), dtype=torch.uint8, device=x.device)])
        return plan.forward(x)[:n]

    @staticmethod
    def infer_batch(model: "CNNModel", batch_data, *, device="cuda") -> np.ndarray:
        """vanilla.py:336-359 -> (N, num_classes) float32 probabilities."""
        return CNNModel.infer_batch_device(model, batch_data, device=device).cpu().numpy()
