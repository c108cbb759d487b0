"""Device plans for the torchvision-ResNet based models: CNNModel (resnet18/34/50 + GAP + Linear + softmax,
vanilla.py:256-333) and UNetModel with the ResNet-50 encoder (unet.py:292-468, `fcn-tissue_mask`).

Post-activation ResNets: every conv is followed by a BatchNorm, which is folded into the packed weights;
the residual add + ReLU run in the epilogue of the block's last conv (fp16 residual, hi+lo in split mode).
3x3 stride-1 convs run as shifted GEMMs on zero-padded NHWC buffers; the 7x7/2 stem and the three
stride-2 3x3 convs (+ their 1x1/2 shortcuts) go through im2col.
"""
from __future__ import annotations

import torch

from tiatoolbox_b200.models import ops
from tiatoolbox_b200.models.architecture._params import bn_scale_shift
from tiatoolbox_b200.models.architecture.resnet_common import BASIC, BOTTLENECK
from tiatoolbox_b200.models.plan import Plan, ceil_to


class ResNetTrunkMixin:
    """Builds stem + layer1..4 into `self` (a Plan).  Returns dense feature Acts [(x0,H0), (x1,H1), ...]."""

    def _trunk(self, sd, prefix, arch, H, W, u8_scale, layer_names=("layer1", "layer2", "layer3", "layer4"), stem_names=("conv1", "bn1"),
               keep_x0=True):
        B = self.B
        bottleneck = arch in BOTTLENECK
        blocks = BOTTLENECK[arch] if bottleneck else BASIC[arch]
        exp = 4 if bottleneck else 1
        # ---- stem: 7x7/2 pad 3 + BN + ReLU (im2col, K = 147 -> 192), then 3x3/2 max pool ----
        H0, W0 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        x0 = self.act("stem.x0", (B, H0, W0), 64)
        self.stem_u8(sd, prefix + stem_names[0], prefix + stem_names[1], inp=self.inp, u8_scale=u8_scale, hw=(H, W), k=7, stride=2, pad=3,
                     out_hw=(H0, W0), N=64, out16=x0)
        H1, W1 = (H0 + 2 - 3) // 2 + 1, (W0 + 2 - 3) // 2 + 1
        feats = [(x0, H0, W0)]
        inplanes = 64
        if bottleneck:
            cur = self.act("pool", (B, H1, W1), 64)
            self.add(lambda cur=cur: ops.maxpool3x3s2(x0, B, H0, W0, 64, cur, H1, W1, 0, 0))
        else:
            cur = self.act("pool.pad", (B, H1 + 2, W1 + 2), 64)
            self.add(lambda cur=cur: ops.maxpool3x3s2(x0, B, H0, W0, 64, cur, H1 + 2, W1 + 2, 1, 1))
        h, w = H1, W1
        for li, (lname, nb) in enumerate(zip(layer_names, blocks)):
            planes = 64 * 2**li
            stride = 1 if li == 0 else 2
            for b in range(nb):
                p = f"{prefix}{lname}.{b}"
                s = stride if b == 0 else 1
                last_of_trunk = li == 3 and b == nb - 1
                if bottleneck:
                    cur, h, w = self._bottleneck(sd, p, cur, h, w, inplanes, planes, s)
                else:
                    cur, h, w = self._basic(sd, p, cur, h, w, inplanes, planes, s, dense_out=last_of_trunk)
                inplanes = planes * exp
            feats.append((cur, h, w))
        return feats

    # ---- torchvision BasicBlock on zero-padded buffers ----
    def _basic(self, sd, p, xin, h, w, cin, planes, s, dense_out):
        B = self.B
        ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
        mid = self.act(p + ".mid", (B, ho + 2, wo + 2), planes)
        if s == 1:
            w1, b1 = self.fold_bn(sd, p + ".conv1", p + ".bn1")
            self.conv(x=xin, grid=(h + 2, w + 2), cin=cin, w=w1, k=3, centered=False, N=planes, window=(0, h, 0, w),
                      dst_geom=(ho + 2, wo + 2, 1, 1), bias=b1, relu=True, out16=mid)
        else:
            w1, b1 = self.fold_bn(sd, p + ".conv1", p + ".bn1", im2col=True)
            i1 = self.act(p + ".im2col", (B * ho * wo,), 9 * cin)
            self.add(lambda: ops.im2col(xin.t, False, 1.0, xin.lo, B, h + 2, w + 2, cin, xin.ld, 3, 3, s, 0, 0, ho, wo, i1.t, i1.ld, 9 * cin, i1.lo))
            self.conv(x=i1, grid=(ho, wo), cin=9 * cin, w=w1, k=1, centered=False, N=planes, window=(0, ho, 0, wo),
                      dst_geom=(ho + 2, wo + 2, 1, 1), bias=b1, relu=True, out16=mid)
        # identity / projection shortcut
        if p + ".downsample.0.weight" in sd:
            wd, bd = self.fold_bn(sd, p + ".downsample.0", p + ".downsample.1")
            g = self.act(p + ".ds_gather", (B * ho * wo,), cin)
            self.add(lambda: ops.im2col(xin.t, False, 1.0, xin.lo, B, h + 2, w + 2, cin, xin.ld, 1, 1, s, -1, -1, ho, wo, g.t, g.ld, cin, g.lo))
            res = self.act(p + ".ds", (B, ho, wo), planes)
            self.gemm(x=g, rows=B * ho * wo, cin=cin, w=wd, N=planes, bias=bd, out16=res)
            res_geom = (ho, wo, 0, 0)
        else:
            res, res_geom = xin, (h + 2, w + 2, 1, 1)
        w2, b2 = self.fold_bn(sd, p + ".conv2", p + ".bn2")
        if dense_out:
            out = self.act(p + ".out", (B, ho, wo), planes)
            dst_geom = (ho, wo, 0, 0)
        else:
            out = self.act(p + ".out.pad", (B, ho + 2, wo + 2), planes)
            dst_geom = (ho + 2, wo + 2, 1, 1)
        self.conv(x=mid, grid=(ho + 2, wo + 2), cin=planes, w=w2, k=3, centered=False, N=planes, window=(0, ho, 0, wo), dst_geom=dst_geom,
                  bias=b2, res16=res, res_geom=res_geom, relu=True, out16=out)
        return out, ho, wo

    # ---- torchvision Bottleneck (v1.5: stride on the 3x3) on dense buffers ----
    def _bottleneck(self, sd, p, xin, h, w, cin, planes, s):
        B = self.B
        ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
        w1, b1 = self.fold_bn(sd, p + ".conv1", p + ".bn1")
        mid1 = self.act(p + ".mid1.pad", (B, h + 2, w + 2), planes)
        self.conv(x=xin, grid=(h, w), cin=cin, w=w1, k=1, centered=False, N=planes, window=(0, h, 0, w), dst_geom=(h + 2, w + 2, 1, 1),
                  bias=b1, relu=True, out16=mid1)
        mid2 = self.act(p + ".mid2", (B, ho, wo), planes)
        if s == 1:
            w2, b2 = self.fold_bn(sd, p + ".conv2", p + ".bn2")
            self.conv(x=mid1, grid=(h + 2, w + 2), cin=planes, w=w2, k=3, centered=False, N=planes, window=(0, h, 0, w),
                      dst_geom=(ho, wo, 0, 0), bias=b2, relu=True, out16=mid2)
        else:
            w2, b2 = self.fold_bn(sd, p + ".conv2", p + ".bn2", im2col=True)
            i2 = self.act(p + ".im2col", (B * ho * wo,), 9 * planes)
            self.add(lambda: ops.im2col(mid1.t, False, 1.0, mid1.lo, B, h + 2, w + 2, planes, mid1.ld, 3, 3, s, 0, 0, ho, wo, i2.t, i2.ld,
                                        9 * planes, i2.lo))
            self.gemm(x=i2, rows=B * ho * wo, cin=9 * planes, w=w2, N=planes, bias=b2, relu=True, out16=mid2)
        if p + ".downsample.0.weight" in sd:
            wd, bd = self.fold_bn(sd, p + ".downsample.0", p + ".downsample.1")
            res = self.act(p + ".ds", (B, ho, wo), planes * 4)
            if s == 1:
                self.gemm(x=xin, rows=B * h * w, cin=cin, w=wd, N=planes * 4, bias=bd, out16=res)
            else:
                g = self.act(p + ".ds_gather", (B * ho * wo,), cin)
                self.add(lambda: ops.im2col(xin.t, False, 1.0, xin.lo, B, h, w, cin, xin.ld, 1, 1, s, 0, 0, ho, wo, g.t, g.ld, cin, g.lo))
                self.gemm(x=g, rows=B * ho * wo, cin=cin, w=wd, N=planes * 4, bias=bd, out16=res)
        else:
            res = xin
        w3, b3 = self.fold_bn(sd, p + ".conv3", p + ".bn3")
        out = self.act(p + ".out", (B, ho, wo), planes * 4)
        self.conv(x=mid2, grid=(ho, wo), cin=planes, w=w3, k=1, centered=False, N=planes * 4, window=(0, ho, 0, wo), bias=b3, res16=res,
                  res_geom=(ho, wo, 0, 0), relu=True, out16=out)
        return out, ho, wo


class CNNPlan(Plan, ResNetTrunkMixin):
    """CNNModel.forward on the device: (B, S, S, 3) uint8 tiles (scaled by `u8_scale`, 1/255 for the kather100k
    preproc, dataset/classification.py:30-63) -> (B, num_classes) softmax probabilities."""

    def __init__(self, sd, *, backbone: str, num_classes: int, batch: int, device, in_size: int = 224, u8_scale: float = 1.0 / 255.0,
                 precision: str = "fp16x3"):
        super().__init__(device, batch, precision)
        S = in_size
        self.inp = self.buf("input_u8", (batch, S, S, 3), torch.uint8)
        feats = self._trunk(sd, "feat_extract.", backbone, S, S, u8_scale, ("4", "5", "6", "7"), ("0", "1"))
        x4, h, w = feats[-1]
        c = x4.C
        self.probs = self.buf("probs", (batch, num_classes), torch.float32)
        wt = self.const(sd["classifier.weight"])
        bb = self.const(sd["classifier.bias"])
        self.add(lambda: ops.gap_fc_softmax(x4, batch, h * w, c, wt, bb, num_classes, None, self.probs))
        self.flops += 2.0 * batch * c * num_classes

    def forward(self, x_u8: torch.Tensor) -> torch.Tensor:
        assert x_u8.shape == self.inp.shape and x_u8.dtype == torch.uint8, (x_u8.shape, self.inp.shape)
        self.inp.copy_(x_u8, non_blocking=True)
        self.run()
        return self.probs


class UNetPlan(Plan, ResNetTrunkMixin):
    """UNetModel(encoder="resnet50", skip_type="add").infer_batch on the device:
    (B, S, S, 3) uint8 -> (B, S/2, S/2, C) probabilities."""

    def __init__(self, sd, *, batch: int, device, in_size: int = 1024, num_output_channels: int = 2, decoder_block=(3,),
                 precision: str = "fp16x3"):
        super().__init__(device, batch, precision)
        if tuple(decoder_block) != (3,):
            msg = "the B200 UNet plan implements decoder_block=[3] (fcn-tissue_mask)"
            raise NotImplementedError(msg)
        B, S = batch, in_size
        self.inp = self.buf("input_u8", (B, S, S, 3), torch.uint8)
        feats = self._trunk(sd, "backbone.", "resnet50", S, S, 1.0 / 255.0)
        x4, h, w = feats[-1]
        x = self.act("dec.x", (B, h, w), 1024)
        self.gemm(x=x4, rows=B * h * w, cin=2048, w=self.weight(sd["conv1x1.weight"]), N=1024, out16=x)
        cur_c = 1024
        down = [1024, 512, 256, 64]
        for i in range(4):
            skip, hs, ws = feats[3 - i]
            assert (hs, ws) == (2 * h, 2 * w) and skip.C == cur_c, (hs, ws, h, w, skip.C, cur_c)
            nxt = down[i + 1] if i + 1 < 4 else down[i]
            s_, t_ = bn_scale_shift(sd, f"uplist.{i}.0")
            sc, sh = self.const(s_), self.const(t_)
            up = self.act(f"dec.up{i}.pad", (B, hs + 2, ws + 2), cur_c)
            self.add(lambda x=x, h=h, w=w, skip=skip, hs=hs, ws=ws, cur_c=cur_c, sc=sc, sh=sh, up=up: ops.upsample_add(
                x, h, w, h, w, 0, 0, skip, hs, ws, 0, 0, B, cur_c, sc, sh, True, up, hs + 2, ws + 2, 1, 1))
            nx = self.act(f"dec.x{i}", (B, hs, ws), nxt)
            self.conv(x=up, grid=(hs + 2, ws + 2), cin=cur_c, w=self.weight(sd[f"uplist.{i}.2.weight"]), k=3, centered=False, N=nxt,
                      window=(0, hs, 0, ws), dst_geom=(hs, ws, 0, 0), out16=nx)
            x, h, w, cur_c = nx, hs, ws, nxt
        self.Ho, self.Wo = h, w
        C = num_output_channels
        self.logits = self.buf("logits", (B, h, w, C), torch.float32)
        wt = self.const(sd["clf.weight"].reshape(C, cur_c))
        bb = self.const(sd["clf.bias"])
        self.add(lambda: ops.pointwise_head(x, B * h * w, cur_c, None, None, False, wt, bb, C, self.logits, C, 0))
        self.flops += 2.0 * B * h * w * cur_c * C
        self.probs = self.buf("probs", (B, h, w, C), torch.float32)
        self.add(lambda: ops.softmax_up2_crop(self.logits, B, h, w, C, self.probs))

    def forward(self, x_u8: torch.Tensor) -> torch.Tensor:
        assert x_u8.shape == self.inp.shape and x_u8.dtype == torch.uint8, (x_u8.shape, self.inp.shape)
        self.inp.copy_(x_u8, non_blocking=True)
        self.run()
        return self.probs
