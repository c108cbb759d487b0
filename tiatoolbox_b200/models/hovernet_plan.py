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
from tiatoolbox_b200.models.architecture.hovernet import RES_CFG, branch_channels, tf_same_pad
from tiatoolbox_b200.models.plan import Plan


class HoVerNetPlan(Plan):
    def __init__(self, sd, *, batch: int, device, mode: str = "fast", num_types: int | None = 6, in_size: int | None = None,
                 precision: str = "fp16x3", num_layers: int | None = None):
        super().__init__(device, batch, precision)
        self.mode, self.num_types, self.num_layers = mode, num_types, num_layers
        S = in_size or (256 if mode == "fast" else 270)
        self.S = S
        B = batch
        k = 5 if mode == "original" else 3
        self.inp = self.buf("input_u8", (B, S, S, 3), torch.uint8)

        # ---------------- conv0: [pad] 7x7 s1 + BN + ReLU via im2col (Cin = 3) ----------------
        pad0 = tf_same_pad(S, 7, 1)[0] if mode == "fast" else 0
        S0 = S if mode == "fast" else S - 6
        # d0's unit-0 conv3 and its 1x1 shortcut are ONE GEMM over the channel concatenation [t2 | x0] (see _res_block), so x0 is
        # produced straight into its slice of that buffer
        name0, cin0, (_, c2_0, _), _, stride0 = RES_CFG[0]
        assert stride0 == 1 and cin0 == 64
        cat0 = self.act(f"{name0}.cat", (B, S0, S0), c2_0 + cin0)
        x0 = cat0.slice(c2_0, cin0)
        self.stem_u8(sd, "conv0./", "conv0.bn", inp=self.inp, u8_scale=1.0 / 255.0, hw=(S, S), k=7, stride=1, pad=pad0,
                     out_hw=(S0, S0), N=64, out16=x0, name="im2col0")

        # ---------------- encoder: 4 pre-activation bottleneck residual blocks ----------------
        feats = []
        x_act, H = x0, S0
        for bi, (name, cin, (c1, c2, c3), units, stride) in enumerate(RES_CFG):
            x_act, H = self._res_block(sd, name, x_act, H, cin, c1, c2, c3, units, stride, cat=cat0 if bi == 0 else None)
            feats.append((x_act, H))
        (d0, H0), (d1, H1), (d2, H2), (d3, H3) = feats
        dbot = self.act("dbot", (B, H3, H3), 1024)
        self.gemm(x=d3, rows=B * H3 * H3, cin=2048, w=self.weight(sd["conv_bot.weight"]), N=1024, out16=dbot)

        # ---------------- decoder ----------------
        G3 = 2 * H3
        assert G3 == H2
        u3in = self.act("u3in", (B, G3, G3), 1024)
        self.add(lambda: ops.upsample_add(dbot, H3, H3, H3, H3, 0, 0, d2, H2, H2, 0, 0, B, 1024, None, None, False, u3in, G3, G3, 0, 0))
        crop0, crop1 = (184, 72) if mode == "original" else (92, 36)
        h = k // 2
        br_ch = branch_channels(num_types, num_layers)   # [tp], np, hv, [ls (HoVerNet+)]
        branches = list(br_ch)
        # The branches have identical shapes, so every decoder layer is ONE grouped launch over R*B images (group r = branch
        # r with its own weight rows): three times the tiles per launch, a third of the launches.
        R = len(branches)
        BR = B * R
        pre = [f"decoder.{b}" for b in branches]
        cb3 = self.act("cb3", (BR, G3, G3), 512)
        t3 = self.act("t3", (BR, G3, G3), 128)
        u3out = self.act("u3out", (BR, G3, G3), 512)
        lo3 = h * 9                      # conva + 8 dense units
        W3 = G3 - 2 * lo3                # valid width of u3's output
        G2 = 2 * W3
        assert G2 == H1 - crop1, (G2, H1, crop1)
        u2in = self.act("u2in", (BR, G2, G2), 512)
        cb2 = self.act("cb2", (BR, G2, G2), 256)
        t2 = self.act("t2d", (BR, G2, G2), 128)
        u2out = self.act("u2out", (BR, G2, G2), 256)
        lo2 = h * 5
        W2 = G2 - 2 * lo2
        So = 2 * W2
        assert So == H0 - crop0, (So, H0, crop0)
        pu = tf_same_pad(So, k, 1)
        G1 = So + pu[0] + pu[1]
        u1in = self.act("u1in_padded", (BR, G1, G1), 256)
        self.So = So

        def bn_cat(names):
            st = [bn_scale_shift(sd, n) for n in names]
            return self.const(torch.cat([a for a, _ in st])), self.const(torch.cat([b_ for _, b_ in st]))

        # u3: conva reads the shared u3in, each branch writes its own block of cb3
        self.conv(x=u3in, grid=(G3, G3), cin=1024, w=self.weight_groups([sd[f"{p}.u3.conva.weight"] for p in pre]), k=k, centered=True,
                  N=256, window=(h, G3 - h, h, G3 - h), out16=cb3, coff16=0, nimg=BR, groups=R, x_grouped=False)
        self._dense(sd, [f"{p}.u3.dense" for p in pre], 8, cb3, t3, G3, 256, k, h)
        sc, sh = bn_cat([f"{p}.u3.dense.blk_bna.bn" for p in pre])
        self.conv(x=cb3, grid=(G3, G3), cin=512, w=self.weight_groups([sd[f"{p}.u3.convf.weight"] for p in pre]), k=1, centered=False,
                  N=512, window=(lo3, G3 - lo3, lo3, G3 - lo3), out16=u3out, pre=(sc, sh, True), nimg=BR, groups=R)
        # u2
        for r in range(R):
            self.add(lambda r=r: ops.upsample_add(u3out.images(r * B, B), W3, W3, G3, G3, lo3, lo3, d1, H1, H1, crop1 // 2, crop1 // 2, B,
                                                  512, None, None, False, u2in.images(r * B, B), G2, G2, 0, 0))
        self.conv(x=u2in, grid=(G2, G2), cin=512, w=self.weight_groups([sd[f"{p}.u2.conva.weight"] for p in pre]), k=k, centered=True,
                  N=128, window=(h, G2 - h, h, G2 - h), out16=cb2, coff16=0, nimg=BR, groups=R)
        self._dense(sd, [f"{p}.u2.dense" for p in pre], 4, cb2, t2, G2, 128, k, h)
        sc, sh = bn_cat([f"{p}.u2.dense.blk_bna.bn" for p in pre])
        self.conv(x=cb2, grid=(G2, G2), cin=256, w=self.weight_groups([sd[f"{p}.u2.convf.weight"] for p in pre]), k=1, centered=False,
                  N=256, window=(lo2, G2 - lo2, lo2, G2 - lo2), out16=u2out, pre=(sc, sh, True), nimg=BR, groups=R)
        # u1
        for r in range(R):
            self.add(lambda r=r: ops.upsample_add(u2out.images(r * B, B), W2, W2, G2, G2, lo2, lo2, d0, H0, H0, crop0 // 2, crop0 // 2, B,
                                                  256, None, None, False, u1in.images(r * B, B), G1, G1, pu[0], pu[0]))
        # u1 conv with u0 (BN + ReLU + 1x1 conv + bias -> fp32 logits) fused into its epilogue: the 64-channel tensor is never stored
        out_chs = [br_ch[b] for b in branches]
        logits = {b: self.buf(f"logits_{b}", (B, So, So, oc), torch.float32) for b, oc in zip(branches, out_chs)}
        w_u1 = self.weight_groups([sd[f"{p}.u1.conva.weight"] for p in pre])
        if max(out_chs) <= 8:
            hw = torch.zeros(R, 8, 64)
            hb = torch.zeros(R, 8)
            for r, (p, oc) in enumerate(zip(pre, out_chs)):
                hw[r, :oc] = sd[f"{p}.u0.conv.weight"].reshape(oc, 64).float()
                hb[r, :oc] = sd[f"{p}.u0.conv.bias"].float()
            hs, ht = bn_cat([f"{p}.u0.bn" for p in pre])
            self.conv(x=u1in, grid=(G1, G1), cin=256, w=w_u1, k=k, centered=False, N=64, window=(0, So, 0, So), dst_geom=(So, So, 0, 0),
                      nimg=BR, groups=R, head=(self.const(hw), self.const(hb), hs, ht, [logits[b] for b in branches]))
            for oc in out_chs:
                self.flops += 2.0 * B * So * So * 64 * oc
        else:   # more than 8 nucleus types: separate head kernels
            u1 = self.act("u1", (BR, So, So), 64)
            self.conv(x=u1in, grid=(G1, G1), cin=256, w=w_u1, k=k, centered=False, N=64, window=(0, So, 0, So), dst_geom=(So, So, 0, 0),
                      out16=u1, nimg=BR, groups=R)
            for r, (b, p, oc) in enumerate(zip(branches, pre, out_chs)):
                s_, t_ = bn_scale_shift(sd, f"{p}.u0.bn")
                sc, sh = self.const(s_), self.const(t_)
                wt = self.const(sd[f"{p}.u0.conv.weight"].reshape(oc, 64))
                bb = self.const(sd[f"{p}.u0.conv.bias"])
                self.add(lambda sc=sc, sh=sh, wt=wt, bb=bb, lg=logits[b], oc=oc, r=r: ops.pointwise_head(
                    u1.images(r * B, B), B * So * So, 64, sc, sh, True, wt, bb, oc, lg, oc, 0))
                self.flops += 2.0 * B * So * So * 64 * oc
        self.logits = logits
        self.np_out = self.buf("np_out", (B, So, So, 1), torch.float32)
        self.hv_out = self.buf("hv_out", (B, So, So, 2), torch.float32)
        self.tp_out = self.buf("tp_out", (B, So, So, 1), torch.float32) if num_types is not None else None
        self.add(lambda: ops.hovernet_heads(logits["np"], logits["hv"], logits.get("tp"), num_types or 0, B * So * So, self.np_out,
                                            self.hv_out, self.tp_out))
        # HoVerNet+: the layer head is argmax(softmax(ls)) as float32 (hovernetplus.py:455-458) — 27k x num_layers values per image,
        # formed with two torch calls after the plan has run (see forward)
        self.ls_out = self.buf("ls_out", (B, So, So, 1), torch.float32) if num_layers is not None else None

    # ------------------------------------------------------------------------------------------
    def _res_block(self, sd, name, x_act, H, cin, c1, c2, c3, units, stride, cat=None):
        """Pre-activation bottleneck block (hovernet.py:152-251).  The residual stream is fp32; unit 0's conv3 and the block's
        1x1 shortcut are a single GEMM over the channel concatenation [t2 | shortcut input] with K-concatenated weights
        [W3 | Wsc], which writes the stream once instead of write + read-modify-write."""
        B = self.B
        Ho = (H + stride - 1) // stride
        stream = self.buf(f"{name}.stream", (B, Ho, Ho, c3), torch.float32)
        out = self.act(f"{name}.out", (B, Ho, Ho), c3)
        ppad = self.act(f"{name}.pad", (B, Ho + 2, Ho + 2), c1)   # TF-same pad of a stride-1 3x3: (1, 1)
        if cat is None:
            cat = self.act(f"{name}.cat", (B, Ho, Ho), c2 + cin)
        t2, sc_in = cat.slice(0, c2), cat.slice(c2, cin)
        # TMA promotes its L2 fetches to 256 bytes: a 64-channel slice (128 B) of the wider concat row would drag its neighbour along on
        # every read (ncu: +0.27 GB per conv3 in d0), so the units that do not need the concatenation get a buffer of their own
        t2_rest = self.act(f"{name}.t2", (B, Ho, Ho), c2) if (c2 * 2 < 256 and units > 1) else t2
        if stride == 1:   # the caller produced x_act straight into the shortcut slice
            assert x_act.t.data_ptr() == sc_in.t.data_ptr() and x_act.ld == sc_in.ld
        else:             # strided 1x1 shortcut: gather the pixels it sees
            self.im2col(x_act.t, False, 1.0, x_act.lo, B, H, H, cin, x_act.ld, 1, stride, 0, Ho, sc_in, cin)
        w3sc = self.weight(torch.cat([sd[f"{name}.units.0.conv3.weight"], sd[f"{name}.shortcut.weight"]], dim=1))
        cur_in, cur_c, cur_pre = x_act, cin, None
        for i in range(units):
            u = f"{name}.units.{i}"
            w1, b1 = self.fold_bn(sd, f"{u}.conv1", f"{u}.conv1/bn")
            s = stride if i == 0 else 1
            Hi = H if i == 0 else Ho
            if s == 1:
                w2, b2 = self.fold_bn(sd, f"{u}.conv2", f"{u}.conv2/bn")
                # conv1 (1x1 + BN + ReLU) straight into the interior of the zero-padded buffer
                self.conv(x=cur_in, grid=(Hi, Hi), cin=cur_c, w=w1, k=1, centered=False, N=c1, window=(0, Hi, 0, Hi),
                          dst_geom=(Hi + 2, Hi + 2, 1, 1), bias=b1, relu=True, out16=ppad, pre=cur_pre)
                self.conv(x=ppad, grid=(Hi + 2, Hi + 2), cin=c1, w=w2, k=3, centered=False, N=c2, window=(0, Hi, 0, Hi),
                          dst_geom=(Hi, Hi, 0, 0), bias=b2, relu=True, out16=t2 if i == 0 else t2_rest)
            else:
                # strided unit: conv1 stores its output space-to-depth (TF-same padding offset included), so the stride-2 3x3 is a
                # stride-1 shifted GEMM over the cell grid -- no im2col
                w2, b2 = self.fold_bn(sd, f"{u}.conv2", f"{u}.conv2/bn")
                pb, _ = tf_same_pad(H, 3, s)
                Hs = Ho + 1
                t1s = self.act(f"{name}.t1_s2d", (B, Hs, Hs), 4 * c1)
                self.conv(x=cur_in, grid=(H, H), cin=cur_c, w=w1, k=1, centered=False, N=c1, window=(0, H, 0, H),
                          dst_geom=(Hs, Hs, pb, pb), bias=b1, relu=True, out16=t1s, s2d_c=c1, pre=cur_pre)
                self.conv_s2d(x=t1s, grid=(Hs, Hs), cin=c1, w=w2, k=3, N=c2, out_hw=(Ho, Ho), bias=b2, relu=True, out16=t2)
            last = i == units - 1
            bn_next = f"{name}.blk_bna.bn" if last else f"{name}.units.{i + 1}.preact/bn"
            sc, sh = bn_scale_shift(sd, bn_next)
            if i == 0:   # conv3 + shortcut in one GEMM, no residual operand
                x3, k3, w3, res = cat, c2 + cin, w3sc, {}
            else:
                x3, k3, w3, res = t2_rest, c2, self.weight(sd[f"{u}.conv3.weight"]), {"res32": stream, "ldres": c3}
            if last:   # closing BN+ReLU of the block: the activated tensor is what the next block / the decoder skip consume
                self.gemm(x=x3, rows=B * Ho * Ho, cin=k3, w=w3, N=c3, out16b=out, scale2=self.const(sc), shift2=self.const(sh), **res)
            else:      # only the fp32 stream is written; the next unit's conv1 applies its pre-activation while loading it
                self.gemm(x=x3, rows=B * Ho * Ho, cin=k3, w=w3, N=c3, out32=stream, ld32=c3, **res)
            cur_in, cur_c, cur_pre = stream, c3, (self.const(sc), self.const(sh), True)
        return out, Ho

    def _dense(self, sd, names, units, cb, t, G, c_in, k, h):
        """Dense blocks (hovernet.py:61-149) of all branches (`names`: one parameter prefix per branch) on one shared grid:
        unit u reads channels [0, c_in + 32u) of `cb` under its own BN+ReLU and appends 32 raw channels; valid windows
        shrink by k//2 per unit.  One grouped launch per layer (group = branch)."""
        R = len(names)
        BR = self.B * R
        for u in range(units):
            uns = [f"{n}.units.{u}" for n in names]
            c_u = c_in + 32 * u
            st = [bn_scale_shift(sd, f"{un}.preact_bna/bn") for un in uns]
            sc = self.const(torch.cat([a for a, _ in st]))
            sh = self.const(torch.cat([b for _, b in st]))
            w1, b1 = self.fold_bn_groups(sd, [f"{un}.conv1" for un in uns], [f"{un}.conv1/bn" for un in uns])
            lo = h * (u + 2)
            # the unit's own BN+ReLU over the raw concat channels is applied while staging the GEMM operand; only the rows /
            # columns the unit's kxk convolution will read are produced (whole tiles outside are skipped)
            self.conv(x=cb, grid=(G, G), cin=c_u, w=w1, k=1, centered=False, N=128, window=(lo - h, G - lo + h, lo - h, G - lo + h),
                      bias=b1, relu=True, out16=t, pre=(sc, sh, True), nimg=BR, groups=R)
            w2 = self.weight_groups([sd[f"{un}.conv2.weight"] for un in uns], groups=4)
            self.conv(x=t, grid=(G, G), cin=128, w=w2, k=k, centered=True, N=32, window=(lo, G - lo, lo, G - lo),
                      out16=cb, coff16=c_u, nimg=BR, groups=R)

    # ------------------------------------------------------------------------------------------
    def forward(self, x_u8: torch.Tensor):
        """x_u8: (B, S, S, 3) uint8 on this plan's device.  Returns (np, hv[, tp]) fp32 device tensors
        (valid until the next call)."""
        assert x_u8.shape == self.inp.shape and x_u8.dtype == torch.uint8, (x_u8.shape, self.inp.shape)
        self.inp.copy_(x_u8, non_blocking=True)
        self.run()
        if self.ls_out is not None:
            probs = torch.softmax(self.logits["ls"], dim=-1)
            self.ls_out.copy_(torch.argmax(probs, dim=-1, keepdim=True))
            return self.np_out, self.hv_out, self.tp_out, self.ls_out
        if self.tp_out is not None:
            return self.np_out, self.hv_out, self.tp_out
        return self.np_out, self.hv_out
