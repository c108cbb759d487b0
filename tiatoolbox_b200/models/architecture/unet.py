"""UNetModel with a ResNet-50 encoder (`fcn-tissue_mask`) for the B200.

Mirrors tiatoolbox/models/architecture/unet.py:292-476 for encoder="resnet50", skip_type="add":
same state_dict keys (`backbone.*` incl. the unused `backbone.fc`, `conv1x1`, `uplist.<i>.{0,2}`,
`clf`, `upsample2x.unpool_mat`), forward (x/255 -> encoder -> 1x1 -> 4 x [up2x + skip -> BN-ReLU-conv3x3]
-> 1x1 clf) and infer_batch tail (softmax -> bilinear x2 -> centre crop)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from tiatoolbox_b200.models.architecture._params import ParamSpec
from tiatoolbox_b200.models.architecture.resnet_common import _bn, resnet_features, resnet_spec


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


def unet_forward_reference(sd, x_nchw: torch.Tensor, decoder_block=(3,)) -> torch.Tensor:
    """fp32 restatement of UNetModel.forward (unet.py:374-418): logits N x C x H/2 x W/2.  x in [0, 255]."""
    en = resnet_features(sd, x_nchw / 255.0, "backbone.", "resnet50")
    x = F.conv2d(en[-1], sd["conv1x1.weight"])
    en = en[:-1]
    for idx in range(1, len(en) + 1):
        y = en[-idx]
        x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) + y
        li = 0
        for ks in decoder_block:
            x = F.relu(_bn(x, sd, f"uplist.{idx - 1}.{li}"))
            x = F.conv2d(x, sd[f"uplist.{idx - 1}.{li + 2}.weight"], padding=(ks - 1) // 2)
            li += 3
    return F.conv2d(x, sd["clf.weight"], sd["clf.bias"])


def unet_infer_tail_reference(logits: torch.Tensor) -> torch.Tensor:
    """UNetModel.infer_batch after the forward (unet.py:458-465): NHWC probabilities of the centre crop."""
    _, _, h, w = logits.shape
    probs = F.softmax(logits, 1)
    probs = F.interpolate(probs, scale_factor=2, mode="bilinear", align_corners=False)
    ch, cw = (2 * h) // 2, (2 * w) // 2  # crop_shape = [H_in // 2, W_in // 2] with H_in = 2h
    t, l_ = (2 * h - ch) // 2, (2 * w - cw) // 2
    return probs[:, :, t:t + ch, l_:l_ + cw].permute(0, 2, 3, 1).contiguous()


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
        sd = {k: v.to(imgs.device) for k, v in self._sd.items()}
        return unet_forward_reference(sd, imgs, self.decoder_block)

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
