"""fp32 PyTorch restatements of the reference networks' forwards — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The floating-point kernels of the conv engine are compared with these (tests/test_*_gpu.py), the seeded test weights are BN-calibrated
with them (oracle/make_golden_cnn.py) and `bench.py`'s baseline legs time them; each is itself pinned to the UNMODIFIED reference
modules (hovernet.py, unet.py, vanilla.py loaded through oracle/ref_shim.py) by the committed goldens under tests/golden/.
Nothing under tiatoolbox_b200/ imports this module: the product's `forward()` / `infer_batch()` run the device plan.

  hovernet_forward_reference  HoVerNet.forward            tiatoolbox/models/architecture/hovernet.py:405-454 (+ hovernetplus.py:87-133)
  hovernet_heads_reference    HoVerNet.infer_batch heads  hovernet.py:897-909
  unet_forward_reference      UNetModel.forward           unet.py:374-418
  unet_infer_tail_reference   UNetModel.infer_batch tail  unet.py:458-465
  cnn_forward_reference       CNNModel.forward            vanilla.py:300-316
  resnet_features             ResNetEncoder._forward_impl unet.py:33-55 / torchvision ResNet trunk
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from tiatoolbox_b200.models.architecture.hovernet import RES_CFG, branch_channels, tf_same_pad
from tiatoolbox_b200.models.architecture.resnet_common import BASIC, BOTTLENECK

# nn.Sequential(*list(model.children())[:-2]) -> indices of the torchvision ResNet children (vanilla.py)
_LAYERS = ("4", "5", "6", "7")
_STEM = ("0", "1")


def _bn(x, sd, name):
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


def _crop_to(x, ref):
    dh, dw = x.shape[2] - ref.shape[2], x.shape[3] - ref.shape[3]
    return x[:, :, dh // 2: x.shape[2] - (dh - dh // 2), dw // 2: x.shape[3] - (dw - dw // 2)]


def _up2(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _residual_block(x, sd, name, units, stride):
    shortcut = F.conv2d(x, sd[f"{name}.shortcut.weight"], stride=stride) if f"{name}.shortcut.weight" in sd else x
    prev = x
    for i in range(units):
        u = f"{name}.units.{i}"
        y = prev
        if i > 0:
            y = F.relu(_bn(y, sd, f"{u}.preact/bn"))
        y = F.relu(_bn(F.conv2d(y, sd[f"{u}.conv1.weight"]), sd, f"{u}.conv1/bn"))
        s = stride if i == 0 else 1
        pb, pa = tf_same_pad(y.shape[2], 3, s)
        y = F.pad(y, (pb, pa, pb, pa))
        y = F.relu(_bn(F.conv2d(y, sd[f"{u}.conv2.weight"], stride=s), sd, f"{u}.conv2/bn"))
        y = F.conv2d(y, sd[f"{u}.conv3.weight"])
        prev = y + shortcut
        shortcut = prev
    return F.relu(_bn(prev, sd, f"{name}.blk_bna.bn"))


def _dense_block(x, sd, name, units):
    prev = x
    for i in range(units):
        u = f"{name}.units.{i}"
        y = F.relu(_bn(prev, sd, f"{u}.preact_bna/bn"))
        y = F.relu(_bn(F.conv2d(y, sd[f"{u}.conv1.weight"]), sd, f"{u}.conv1/bn"))
        y = F.conv2d(y, sd[f"{u}.conv2.weight"], groups=4)
        prev = torch.cat([_crop_to(prev, y), y], dim=1)
    return F.relu(_bn(prev, sd, f"{name}.blk_bna.bn"))


def hovernet_forward_reference(sd, x_nchw: torch.Tensor, mode: str, num_types: int | None, capture: dict | None = None,
                               num_layers: int | None = None) -> OrderedDict:
    """fp32 PyTorch restatement of HoVerNet.forward (hovernet.py:405-454; with `num_layers` HoVerNetPlus, which only adds the `ls`
    decoder branch: hovernetplus.py:87-133).  x in [0, 255]."""
    x = x_nchw / 255.0
    if mode == "fast":
        pb, pa = tf_same_pad(x.shape[2], 7, 1)
        x = F.pad(x, (pb, pa, pb, pa))
    d = F.relu(_bn(F.conv2d(x, sd["conv0./.weight"]), sd, "conv0.bn"))
    if capture is not None:
        capture["x0"] = d
    feats = []
    for name, _, _, units, stride in RES_CFG:
        d = _residual_block(d, sd, name, units, stride)
        feats.append(d)
        if capture is not None:
            capture[name] = d
    feats[3] = F.conv2d(feats[3], sd["conv_bot.weight"])
    if capture is not None:
        capture["dbot"] = feats[3]
    c0, c1 = (184, 72) if mode == "original" else (92, 36)
    d0 = feats[0][:, :, c0 // 2: feats[0].shape[2] - c0 // 2, c0 // 2: feats[0].shape[3] - c0 // 2]
    d1 = feats[1][:, :, c1 // 2: feats[1].shape[2] - c1 // 2, c1 // 2: feats[1].shape[3] - c1 // 2]
    k = 5 if mode == "original" else 3
    out = OrderedDict()
    for b in branch_channels(num_types, num_layers):
        p = f"decoder.{b}"
        u3 = _up2(feats[3]) + feats[2]
        if capture is not None:
            capture["u3in"] = u3
        u3 = F.conv2d(u3, sd[f"{p}.u3.conva.weight"])
        if capture is not None:
            capture[f"{b}.u3.conva"] = u3
        u3 = F.conv2d(_dense_block(u3, sd, f"{p}.u3.dense", 8), sd[f"{p}.u3.convf.weight"])
        if capture is not None:
            capture[f"{b}.u3out"] = u3
        u2 = _up2(u3) + d1
        if capture is not None:
            capture[f"{b}.u2in"] = u2
        u2 = F.conv2d(u2, sd[f"{p}.u2.conva.weight"])
        u2 = F.conv2d(_dense_block(u2, sd, f"{p}.u2.dense", 4), sd[f"{p}.u2.convf.weight"])
        if capture is not None:
            capture[f"{b}.u2out"] = u2
        u1 = _up2(u2) + d0
        if capture is not None:
            capture[f"{b}.u1in"] = u1
        pb, pa = tf_same_pad(u1.shape[2], k, 1)
        u1 = F.conv2d(F.pad(u1, (pb, pa, pb, pa)), sd[f"{p}.u1.conva.weight"])
        if capture is not None:
            capture[f"{b}.u1"] = u1
        u0 = F.relu(_bn(u1, sd, f"{p}.u0.bn"))
        out[b] = F.conv2d(u0, sd[f"{p}.u0.conv.weight"], sd[f"{p}.u0.conv.bias"])
    return out


def hovernet_heads_reference(out: OrderedDict) -> tuple[torch.Tensor, ...]:
    """Head activations of HoVerNet.infer_batch (hovernet.py:897-909): NHWC, np = softmax[..., 1:],
    tp = argmax as float32, hv raw.  Returns (np, hv[, tp])."""
    o = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in out.items()}
    o["np"] = F.softmax(o["np"], dim=-1)[..., 1:]
    if "tp" in o:
        tp = torch.argmax(F.softmax(o["tp"], dim=-1), dim=-1, keepdim=True).type(torch.float32)
        if "ls" in o:   # HoVerNetPlus.infer_batch (hovernetplus.py:444-459)
            ls = torch.argmax(F.softmax(o["ls"], dim=-1), dim=-1, keepdim=True).type(torch.float32)
            return o["np"], o["hv"], tp, ls
        return o["np"], o["hv"], tp
    return o["np"], o["hv"]


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


def cnn_forward_reference(sd, x_nchw: torch.Tensor, backbone: str) -> torch.Tensor:
    """fp32 restatement of CNNModel.forward (vanilla.py:300-316); x is what the dataset feeds (e.g. [0,1])."""
    feat = resnet_features(sd, x_nchw, "feat_extract.", backbone, _LAYERS, _STEM)[-1]
    gap = feat.mean(dim=(2, 3))
    return torch.softmax(F.linear(gap, sd["classifier.weight"], sd["classifier.bias"]), -1)
