"""UNetModel with a ResNet-50 encoder (`fcn-tissue_mask`) for the B200.

Mirrors tiatoolbox/models/architecture/unet.py:292-476 for encoder="resnet50", skip_type="add":
same state_dict keys (`backbone.*` incl. the unused `backbone.fc`, `conv1x1`, `uplist.<i>.{0,2}`,
`clf`, `upsample2x.unpool_mat`), forward (x/255 -> encoder -> 1x1 -> 4 x [up2x + skip -> BN-ReLU-conv3x3]
-> 1x1 clf) and infer_batch tail (softmax -> bilinear x2 -> centre crop)."""
from __future__ import annotations

import torch

from tiatoolbox_b200.models.architecture._params import ParamSpec
from tiatoolbox_b200.models.architecture.resnet_common import resnet_spec


def unet_spec(num_input_channels: int = 3, num_output_channels: int = 2, decoder_block=(3,)) -> ParamSpec:
    sp = ParamSpec()
    resnet_spec(sp, "backbone.", "resnet50", num_input_channels, with_fc=True)
    down = [2048, 1024, 512, 256, 64]
    sp.conv("conv1x1", down[1], down[0], 1)
    nxt = None
    for i, ch in enumerate(down[1:]):
        nxt = down[i + 2] if i + 2 < len(down) else ch
        cin = ch
        li = 0
        for ks in decoder_block:
            sp.bn(f"uplist.{i}.{li}", cin)
            sp.conv(f"uplist.{i}.{li + 2}", nxt, cin, ks)
            cin = nxt
            li += 3
    sp.conv("clf", num_output_channels, nxt, 1, bias=True)
    sp["upsample2x.unpool_mat"] = ((2, 2), "ones")
    return sp


# =================================================================================================
# Model class (API mirror of tiatoolbox/models/architecture/unet.py:231-476)
# =================================================================================================
import numpy as np  # noqa: E402

from tiatoolbox_b200.models.architecture._params import seeded_state_dict  # noqa: E402
from tiatoolbox_b200.models.models_abc import ModelABC  # noqa: E402


class UNetModel(ModelABC):
    """U-Net with a ResNet-50 encoder (`fcn-tissue_mask`)."""

    def __init__(self, num_input_channels: int = 2, num_output_channels: int = 2, encoder: str = "resnet50",
                 encoder_levels: list[int] | None = None, decoder_block: tuple[int] | None = None, skip_type: str = "add") -> None:
        super().__init__()
        if encoder.lower() not in {"resnet50", "unet"}:
            msg = f"Unknown encoder `{encoder}`"
            raise ValueError(msg)
        if skip_type.lower() not in {"add", "concat"}:
            msg = f"Unknown type of skip connection: `{skip_type}`"
            raise ValueError(msg)
        if encoder.lower() != "resnet50" or skip_type.lower() != "add" or num_input_channels != 3:
            msg = "the B200 path implements UNetModel(3, C, encoder='resnet50', skip_type='add') (fcn-tissue_mask)"
            raise NotImplementedError(msg)
        self.decoder_block = tuple(decoder_block) if decoder_block is not None else (3, 3)
        self.num_output_channels = num_output_channels
        self.skip_type = skip_type.lower()
        self._sd = seeded_state_dict(unet_spec(num_input_channels, num_output_channels, self.decoder_block), seed=0)

    def forward(self, imgs: torch.Tensor, *args, **kwargs) -> torch.Tensor:  # noqa: ARG002
        """unet.py:374-418 on the device plan: NCHW input holding uint8 pixel values in [0, 255] -> logits (N, C, H/2, W/2)."""
        x = ModelABC._nchw_to_u8_nhwc(imgs, self._device)
        plan = self.plan_for(x.shape[0], self._device, x.shape[1:3])
        plan.forward(x)
        return plan.logits.permute(0, 3, 1, 2).clone()

    def _build_plan(self, batch, device, shape):
        from tiatoolbox_b200.models.resnet_plan import UNetPlan

        return UNetPlan(self._sd, batch=batch, device=device, in_size=int(shape[0]), num_output_channels=self.num_output_channels,
                        decoder_block=self.decoder_block, precision=self.precision)

    def postproc(self, image: np.ndarray) -> np.ndarray:
        """argmax along the last axis (unet.py:470-476)."""
        return np.argmax(image, axis=-1)

    @staticmethod
    def infer_batch_device(model: "UNetModel", batch_data, *, device) -> torch.Tensor:
        x = ModelABC._as_device_u8(batch_data, device)
        n = x.shape[0]
        plan_b = getattr(model, "_plan_batch", None) or n
        plan = model.plan_for(plan_b, device, x.shape[1:3])
        if n < plan_b:
            x = torch.cat([x, torch.zeros((plan_b - n, *x.shape[1:]), dtype=torch.uint8, device=x.device)])
        return plan.forward(x)[:n]

    @staticmethod
    def infer_batch(model: "UNetModel", batch_data, *, device) -> np.ndarray:
        """unet.py:420-468 -> (N, H/2, W/2, C) float32 probabilities."""
        return UNetModel.infer_batch_device(model, batch_data, device=device).cpu().numpy()
