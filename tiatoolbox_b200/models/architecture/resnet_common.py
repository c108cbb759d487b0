"""torchvision-style ResNet parameter specs + fp32 reference forward (restated from the published
architecture, He et al. 2016; torchvision.models.resnet v1.5: stride on the 3x3 of a bottleneck).
Used by CNNModel (vanilla.py) and the ResNet-50 U-Net encoder (unet.py)."""
from __future__ import annotations

import torch.nn.functional as F

from tiatoolbox_b200.models.architecture._params import ParamSpec

BASIC = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}
BOTTLENECK = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


def resnet_spec(sp: ParamSpec, prefix: str, arch: str, in_ch: int = 3, layer_names=("layer1", "layer2", "layer3", "layer4"),
                stem_names=("conv1", "bn1"), with_fc: bool = False) -> int:
    """Append a ResNet trunk's parameters under `prefix`; returns the number of output channels."""
    bottleneck = arch in BOTTLENECK
    blocks = BOTTLENECK[arch] if bottleneck else BASIC[arch]
    exp = 4 if bottleneck else 1
    sp.conv(prefix + stem_names[0], 64, in_ch, 7)
    sp.bn(prefix + stem_names[1], 64)
    inplanes = 64
    for li, (lname, nb) in enumerate(zip(layer_names, blocks)):
        planes = 64 * 2**li
        stride = 1 if li == 0 else 2
        for b in range(nb):
            p = f"{prefix}{lname}.{b}"
            s = stride if b == 0 else 1
            if bottleneck:
                sp.conv(p + ".conv1", planes, inplanes, 1)
                sp.bn(p + ".bn1", planes)
                sp.conv(p + ".conv2", planes, planes, 3)
                sp.bn(p + ".bn2", planes)
                sp.conv(p + ".conv3", planes * 4, planes, 1)
                sp.bn(p + ".bn3", planes * 4, gain=0.25)
            else:
                sp.conv(p + ".conv1", planes, inplanes, 3)
                sp.bn(p + ".bn1", planes)
                sp.conv(p + ".conv2", planes, planes, 3)
                sp.bn(p + ".bn2", planes, gain=0.25)
            if s != 1 or inplanes != planes * exp:
                sp.conv(p + ".downsample.0", planes * exp, inplanes, 1)
                sp.bn(p + ".downsample.1", planes * exp)
            inplanes = planes * exp
    if with_fc:
        sp.linear(prefix + "fc", 1000, inplanes)
    return inplanes


def _bn(x, sd, name, calibrate_gain: float = 1.0):
    if sd.get("__calibrate__", False):
        sd[name + ".running_mean"] = x.mean(dim=(0, 2, 3)).detach()
        sd[name + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).detach() + 1e-3
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        False, 0.0, 1e-5)


def resnet_features(sd, x, prefix: str, arch: str, layer_names=("layer1", "layer2", "layer3", "layer4"), stem_names=("conv1", "bn1")):
    """Returns [x0 (after stem relu), x1..x4] like ResNetEncoder._forward_impl (unet.py:33-55)."""
    bottleneck = arch in BOTTLENECK
    blocks = BOTTLENECK[arch] if bottleneck else BASIC[arch]
    x = F.relu(_bn(F.conv2d(x, sd[prefix + stem_names[0] + ".weight"], stride=2, padding=3), sd, prefix + stem_names[1]))
    feats = [x]
    x = F.max_pool2d(x, 3, 2, 1)
    for li, (lname, nb) in enumerate(zip(layer_names, blocks)):
        stride = 1 if li == 0 else 2
        for b in range(nb):
            p = f"{prefix}{lname}.{b}"
            s = stride if b == 0 else 1
            idt = x
            if bottleneck:
                y = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
                y = F.relu(_bn(F.conv2d(y, sd[p + ".conv2.weight"], stride=s, padding=1), sd, p + ".bn2"))
                y = _bn(F.conv2d(y, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            else:
                y = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], stride=s, padding=1), sd, p + ".bn1"))
                y = _bn(F.conv2d(y, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2")
            if p + ".downsample.0.weight" in sd:
                idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=s), sd, p + ".downsample.1")
            x = F.relu(y + idt)
        feats.append(x)
    return feats
