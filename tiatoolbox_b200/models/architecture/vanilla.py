"""CNNModel (torchvision backbone + linear classifier) for the B200 — `resnet18-kather100k` and friends.

Mirrors tiatoolbox/models/architecture/vanilla.py:256-359: same constructor, state_dict keys
(`feat_extract.<i>...`, `classifier.*`), forward semantics (trunk -> GAP -> Linear -> softmax) and
`infer_batch` / `postproc` contract.  Only the ResNet family is on the B200 path."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from tiatoolbox_b200.models.architecture._params import ParamSpec
from tiatoolbox_b200.models.architecture.resnet_common import BASIC, BOTTLENECK, resnet_features, resnet_spec

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


def cnn_forward_reference(sd, x_nchw: torch.Tensor, backbone: str) -> torch.Tensor:
    """fp32 restatement of CNNModel.forward (vanilla.py:300-316); x is what the dataset feeds (e.g. [0,1])."""
    feat = resnet_features(sd, x_nchw, "feat_extract.", backbone, _LAYERS, _STEM)[-1]
    gap = feat.mean(dim=(2, 3))
    return torch.softmax(F.linear(gap, sd["classifier.weight"], sd["classifier.bias"]), -1)
