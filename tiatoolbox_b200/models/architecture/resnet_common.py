"""torchvision-style ResNet parameter specs + fp32 reference forward (restated from the published
architecture, He et al. 2016; torchvision.models.resnet v1.5: stride on the 3x3 of a bottleneck).
Used by CNNModel (vanilla.py) and the ResNet-50 U-Net encoder (unet.py)."""
from __future__ import annotations


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
