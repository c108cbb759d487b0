"""HoVer-Net forward as a device plan (tcgen05 shifted-GEMM convs + a few memory-bound glue kernels).

Follows the layer structure of tiatoolbox/models/architecture/hovernet.py:330-500 (see
architecture/hovernet.py in this package for the fp32 restatement used as the parity reference).
All decoder tensors of one level share ONE coordinate grid: `valid` 3x3/5x5 convolutions use taps
centred on the output pixel, so the reference's centre-crops (hovernet.py:143-149, 436-438) are
just shrinking validity windows, and the dense blocks' channel concatenation is a channel slice of
one wide NHWC buffer.
"""
from __future__ import annotations

import torch

from tiatoolbox_b200.models import ops
from tiatoolbox_b200.models.architecture._params import bn_scale_shift
from tiatoolbox_b200.models.architecture.hovernet import RES_CFG, tf_same_pad
from tiatoolbox_b200.models.plan import Plan, ceil_to, fold_bn, pack_weight


class HoVerNetPlan(Plan):
    def __init__(self, sd, *, batch: int, device, mode: str = "fast", num_types: int | None = 6, in_size: int | None = None):
        super().__init__(device, batch)
        self.mode, self.num_types = mode, num_types
        S = in_size or (256 if mode == "fast" else 270)
        self.S = S
        B, dev = batch, self.dev
        k = 5 if mode == "original" else 3
        self.inp = self.buf("input_u8", (B, S, S, 3), torch.uint8)

        # ---------------- conv0: [pad] 7x7 s1 + BN + ReLU via im2col (Cin = 3) ----------------
        pad0 = tf_same_pad(S, 7, 1)[0] if mode == "fast" else 0
        S0 = S if mode == "fast" else S - 6
        scale0, shift0 = bn_scale_shift(sd, "conv0.bn")
        w0 = (sd["conv0./.weight"].double() * scale0.double().view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(64, 147)
        w0p = torch.zeros(64, 192, dtype=torch.float64)
        w0p[:, :147] = w0
        w0p = self.const(w0p, torch.float16)
        i0 = self.buf("im2col0", (B * S0 * S0, 192))
        x0 = self.buf("x0", (B, S0, S0, 64))
        self.add(lambda: ops.im2col(self.inp, True, 1.0 / 255.0, B, S, S, 3, 3, 7, 7, 1, pad0, pad0, S0, S0, i0, 192))
        self.gemm(x=i0, rows=B * S0 * S0, cin=192, w=w0p, N=64, bias=self.const(shift0), relu=True, out16=x0, ld16=64)

        # ---------------- encoder: 4 pre-activation bottleneck residual blocks ----------------
        feats = []
        x_act, H = x0, S0
        for name, cin, (c1, c2, c3), units, stride in RES_CFG:
            x_act, H = self._res_block(sd, name, x_act, H, cin, c1, c2, c3, units, stride)
            feats.append((x_act, H))
        (d0, H0), (d1, H1), (d2, H2), (d3, H3) = feats
        wbot = pack_weight(sd["conv_bot.weight"], device=dev)
        dbot = self.buf("dbot", (B, H3, H3, 1024))
        self.gemm(x=d3, rows=B * H3 * H3, cin=2048, w=wbot, N=1024, out16=dbot, ld16=1024)

        # ---------------- decoder ----------------
        G3 = 2 * H3
        assert G3 == H2
        u3in = self.buf("u3in", (B, G3, G3, 1024))
        self.add(lambda: ops.upsample_add(dbot, H3, H3, H3, H3, 0, 0, d2, H2, H2, 0, 0, B, 1024, None, None, False, u3in, G3, G3, 0, 0))
        crop0, crop1 = (184, 72) if mode == "original" else (92, 36)
        h = k // 2
        branches = (["tp"] if num_types is not None else []) + ["np", "hv"]
        # scratch shared by the branches (they run back to back on one stream)
        cb3 = self.buf("cb3", (B, G3, G3, 512))
        act3 = self.buf("act3", (B * G3 * G3, 512))
        t3 = self.buf("t3", (B, G3, G3, 128))
        u3out = self.buf("u3out", (B, G3, G3, 512))
        lo3 = h * 9                      # conva + 8 dense units
        W3 = G3 - 2 * lo3                # valid width of u3's output
        G2 = 2 * W3
        assert G2 == H1 - crop1, (G2, H1, crop1)
        u2in = self.buf("u2in", (B, G2, G2, 512))
        cb2 = self.buf("cb2", (B, G2, G2, 256))
        act2 = self.buf("act2", (B * G2 * G2, 256))
        t2 = self.buf("t2d", (B, G2, G2, 128))
        u2out = self.buf("u2out", (B, G2, G2, 256))
        lo2 = h * 5
        W2 = G2 - 2 * lo2
        So = 2 * W2
        assert So == H0 - crop0, (So, H0, crop0)
        pu = tf_same_pad(So, k, 1)
        G1 = So + pu[0] + pu[1]
        u1in = self.buf("u1in_padded", (B, G1, G1, 256))
        u1 = self.buf("u1", (B, So, So, 64))
        self.So = So
        logits = {}
        for b in branches:
            p = f"decoder.{b}"
            out_ch = num_types if b == "tp" else 2
            # u3
            self.conv(x=u3in, grid=(G3, G3), cin=1024, w=pack_weight(sd[f"{p}.u3.conva.weight"], device=dev), k=k, centered=True,
                      N=256, window=(h, G3 - h, h, G3 - h), out16=cb3, ld16=512, coff16=0)
            self._dense(sd, f"{p}.u3.dense", 8, cb3, act3, t3, G3, 256, 512, k, h)
            s, t = bn_scale_shift(sd, f"{p}.u3.dense.blk_bna.bn")
            sc, sh = self.const(s), self.const(t)
            self.add(lambda sc=sc, sh=sh: ops.bn_relu(cb3, False, 512, 0, sc, sh, True, B * G3 * G3, 512, act3, 512, 0))
            self.gemm(x=act3, rows=B * G3 * G3, cin=512, w=pack_weight(sd[f"{p}.u3.convf.weight"], device=dev), N=512,
                      out16=u3out, ld16=512)
            # u2
            self.add(lambda: ops.upsample_add(u3out, W3, W3, G3, G3, lo3, lo3, d1, H1, H1, crop1 // 2, crop1 // 2, B, 512, None, None,
                                              False, u2in, G2, G2, 0, 0))
            self.conv(x=u2in, grid=(G2, G2), cin=512, w=pack_weight(sd[f"{p}.u2.conva.weight"], device=dev), k=k, centered=True,
                      N=128, window=(h, G2 - h, h, G2 - h), out16=cb2, ld16=256, coff16=0)
            self._dense(sd, f"{p}.u2.dense", 4, cb2, act2, t2, G2, 128, 256, k, h)
            s, t = bn_scale_shift(sd, f"{p}.u2.dense.blk_bna.bn")
            sc, sh = self.const(s), self.const(t)
            self.add(lambda sc=sc, sh=sh: ops.bn_relu(cb2, False, 256, 0, sc, sh, True, B * G2 * G2, 256, act2, 256, 0))
            self.gemm(x=act2, rows=B * G2 * G2, cin=256, w=pack_weight(sd[f"{p}.u2.convf.weight"], device=dev), N=256,
                      out16=u2out, ld16=256)
            # u1
            self.add(lambda: ops.upsample_add(u2out, W2, W2, G2, G2, lo2, lo2, d0, H0, H0, crop0 // 2, crop0 // 2, B, 256, None, None,
                                              False, u1in, G1, G1, pu[0], pu[0]))
            self.conv(x=u1in, grid=(G1, G1), cin=256, w=pack_weight(sd[f"{p}.u1.conva.weight"], device=dev), k=k, centered=False,
                      N=64, window=(0, So, 0, So), dst_geom=(So, So, 0, 0), out16=u1, ld16=64)
            # u0: BN + ReLU + 1x1 (+bias) -> fp32 logits
            s, t = bn_scale_shift(sd, f"{p}.u0.bn")
            sc, sh = self.const(s), self.const(t)
            wt = self.const(sd[f"{p}.u0.conv.weight"].reshape(out_ch, 64))
            bb = self.const(sd[f"{p}.u0.conv.bias"])
            lg = self.buf(f"logits_{b}", (B, So, So, out_ch), torch.float32)
            logits[b] = lg
            self.add(lambda sc=sc, sh=sh, wt=wt, bb=bb, lg=lg, out_ch=out_ch: ops.pointwise_head(
                u1, B * So * So, 64, 64, sc, sh, True, wt, bb, out_ch, lg, out_ch, 0))
            self.flops += 2.0 * B * So * So * 64 * out_ch
        self.logits = logits
        self.np_out = self.buf("np_out", (B, So, So, 1), torch.float32)
        self.hv_out = self.buf("hv_out", (B, So, So, 2), torch.float32)
        self.tp_out = self.buf("tp_out", (B, So, So, 1), torch.float32) if num_types is not None else None
        self.add(lambda: ops.hovernet_heads(logits["np"], logits["hv"], logits.get("tp"), num_types or 0, B * So * So, self.np_out,
                                            self.hv_out, self.tp_out))

    # ------------------------------------------------------------------------------------------
    def _res_block(self, sd, name, x_act, H, cin, c1, c2, c3, units, stride):
        B, dev = self.B, self.dev
        Ho = (H + stride - 1) // stride
        stream = self.buf(f"{name}.stream", (B, Ho, Ho, c3), torch.float32)
        act = self.buf(f"{name}.act", (B, Ho, Ho, c3)) if units > 1 else None
        out = self.buf(f"{name}.out", (B, Ho, Ho, c3))
        ppad = self.buf(f"{name}.pad", (B, Ho + 2, Ho + 2, c1))   # TF-same pad of a stride-1 3x3: (1, 1)
        t2 = self.buf(f"{name}.t2", (B, Ho, Ho, c2))
        # ---- shortcut of unit 0 -> fp32 residual stream ----
        wsc = pack_weight(sd[f"{name}.shortcut.weight"], device=dev)
        if stride == 1:
            self.gemm(x=x_act, rows=B * H * H, cin=cin, w=wsc, N=c3, out32=stream, ld32=c3)
        else:
            g = self.buf(f"{name}.sc_gather", (B * Ho * Ho, cin))
            self.add(lambda: ops.im2col(x_act, False, 1.0, B, H, H, cin, cin, 1, 1, stride, 0, 0, Ho, Ho, g, cin))
            self.gemm(x=g, rows=B * Ho * Ho, cin=cin, w=wsc, N=c3, out32=stream, ld32=c3)
        cur_in, cur_c = x_act, cin
        for i in range(units):
            u = f"{name}.units.{i}"
            w1, b1 = fold_bn(sd, f"{u}.conv1", f"{u}.conv1/bn", device=dev)
            w2, b2 = fold_bn(sd, f"{u}.conv2", f"{u}.conv2/bn", device=dev)
            w3 = pack_weight(sd[f"{u}.conv3.weight"], device=dev)
            s = stride if i == 0 else 1
            Hi = H if i == 0 else Ho
            if s == 1:
                # conv1 (1x1 + BN + ReLU) straight into the interior of the zero-padded buffer
                self.conv(x=cur_in, grid=(Hi, Hi), cin=cur_c, w=w1, k=1, centered=False, N=c1, window=(0, Hi, 0, Hi),
                          dst_geom=(Hi + 2, Hi + 2, 1, 1), bias=b1, relu=True, out16=ppad, ld16=c1)
                self.conv(x=ppad, grid=(Hi + 2, Hi + 2), cin=c1, w=w2, k=3, centered=False, N=c2, window=(0, Hi, 0, Hi),
                          dst_geom=(Hi, Hi, 0, 0), bias=b2, relu=True, out16=t2, ld16=c2)
            else:
                t1 = self.buf(f"{name}.t1_full", (B, H, H, c1))
                self.gemm(x=cur_in, rows=B * H * H, cin=cur_c, w=w1, N=c1, bias=b1, relu=True, out16=t1, ld16=c1)
                pb, _ = tf_same_pad(H, 3, s)
                i2 = self.buf(f"{name}.im2col", (B * Ho * Ho, 9 * c1))
                self.add(lambda t1=t1, pb=pb, i2=i2, s=s: ops.im2col(t1, False, 1.0, B, H, H, c1, c1, 3, 3, s, pb, pb, Ho, Ho, i2, 9 * c1))
                self.gemm(x=i2, rows=B * Ho * Ho, cin=9 * c1, w=w2, N=c2, bias=b2, relu=True, out16=t2, ld16=c2)
            last = i == units - 1
            bn_next = f"{name}.blk_bna.bn" if last else f"{name}.units.{i + 1}.preact/bn"
            sc, sh = bn_scale_shift(sd, bn_next)
            self.gemm(x=t2, rows=B * Ho * Ho, cin=c2, w=w3, N=c3, res32=stream, ldres=c3,
                      out32=None if last else stream, ld32=c3, out16b=out if last else act, ld16b=c3,
                      scale2=self.const(sc), shift2=self.const(sh))
            cur_in, cur_c = act, c3
        return out, Ho

    def _dense(self, sd, name, units, cb, act, t, G, c_in, ld, k, h):
        """Dense block (hovernet.py:61-149) on one shared grid: unit u reads channels [0, c_in + 32u) of `cb`
        under its own BN+ReLU and appends 32 raw channels; valid windows shrink by k//2 per unit."""
        B, dev = self.B, self.dev
        for u in range(units):
            un = f"{name}.units.{u}"
            c_u = c_in + 32 * u
            s, sft = bn_scale_shift(sd, f"{un}.preact_bna/bn")
            sc, sh = self.const(s), self.const(sft)
            self.add(lambda sc=sc, sh=sh, c_u=c_u: ops.bn_relu(cb, False, ld, 0, sc, sh, True, B * G * G, c_u, act, ld, 0))
            w1, b1 = fold_bn(sd, f"{un}.conv1", f"{un}.conv1/bn", device=dev)
            self.gemm(x=act, rows=B * G * G, cin=c_u, x_ld=ld, w=w1, N=128, bias=b1, relu=True, out16=t, ld16=128)
            lo = h * (u + 2)
            w2 = pack_weight(sd[f"{un}.conv2.weight"], groups=4, device=dev)
            self.conv(x=t, grid=(G, G), cin=128, w=w2, k=k, centered=True, N=32, window=(lo, G - lo, lo, G - lo),
                      out16=cb, ld16=ld, coff16=c_u)

    # ------------------------------------------------------------------------------------------
    def forward(self, x_u8: torch.Tensor):
        """x_u8: (B, S, S, 3) uint8 on this plan's device.  Returns (np, hv[, tp]) fp32 device tensors
        (valid until the next call)."""
        assert x_u8.shape == self.inp.shape and x_u8.dtype == torch.uint8, (x_u8.shape, self.inp.shape)
        self.inp.copy_(x_u8, non_blocking=True)
        self.run()
        if self.tp_out is not None:
            return self.np_out, self.hv_out, self.tp_out
        return self.np_out, self.hv_out
