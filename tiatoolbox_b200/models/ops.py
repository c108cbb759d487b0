"""Thin torch-tensor wrappers over the convolution-engine entry points of libtiab200.so.

Everything here is plumbing (pointers, pitches, geometry); the arithmetic is in csrc/conv.cu.
Tensors are NHWC and live on one CUDA device; launches go to torch's current stream so that the
whole network can be captured in a CUDA graph.

`Act` describes an fp16 activation tensor: `C` channels per plane, row pitch `ld`, and an optional
second plane at channel offset `lo` holding x - fp16(x) (split-precision mode; lo = 0 -> absent).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from tiatoolbox_b200 import _lib


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


@dataclass
class Act:
    t: torch.Tensor
    C: int
    ld: int
    lo: int = 0

    @property
    def dev(self):
        return self.t.device

    def hi(self) -> torch.Tensor:
        return self.t[..., : self.C]

    def images(self, i0: int, n: int) -> "Act":
        """Images [i0, i0 + n) of the leading (batch) dimension."""
        return Act(self.t[i0:i0 + n], self.C, self.ld, self.lo)

    def slice(self, c0: int, C: int) -> "Act":
        """Channels [c0, c0 + C) of every plane as an Act of its own (same pitch and lo offset, shifted base pointer)."""
        assert c0 % 8 == 0 and c0 + C <= self.C
        return Act(self.t[..., c0:], C, self.ld, self.lo)

    def value(self) -> torch.Tensor:
        """fp32 value (hi + lo) — for tests/debugging."""
        v = self.t[..., : self.C].float()
        if self.lo:
            v = v + self.t[..., self.lo: self.lo + self.C].float()
        return v


class ConvOp:
    """One shifted-GEMM convolution launch (see tiab_conv_desc_t)."""

    def __init__(self, *, x: torch.Tensor, x_rows: int, x_ld: int, cin: int, w: torch.Tensor, taps: list[int],
                 M: int, N: int, grid: tuple[int, int], window: tuple[int, int, int, int],
                 dst_geom: tuple[int, int, int, int], x_lo_off: int = 0, out16=None, ld16=0, coff16=0, lo16=0, out32=None,
                 ld32=0, coff32=0, out16b=None, ld16b=0, coff16b=0, lo16b=0, scale2=None, shift2=None, bias=None,
                 res32=None, res16=None, ldres=0, coffres=0, reslo=0, res_geom=(1, 1, 0, 0), relu=False, pre_scale=None,
                 pre_shift=None, pre_relu=False, x_is_f32=False, split=False, groups: int = 1, group_rows: int = 0,
                 x_grouped: bool = True, tap_coff: list[int] | None = None, s2d_c: int = 0, head=None):
        """`head` = (w [G,8,64] fp32, b [G,8], scale [G,64], shift [G,64], outs: list of fp32 [rows_g, nout_g] tensors): fused
        BN + ReLU + 1x1 conv (+bias) on the 64 output channels, one output tensor per group."""
        d = _lib.ConvDesc()
        d.x, d.x_rows, d.x_ld, d.cin, d.x_lo_off = x.data_ptr(), x_rows, x_ld, cin, x_lo_off
        assert w.dtype == torch.float16 and w.dim() == 2 and w.is_contiguous()
        d.w, d.w_ld = w.data_ptr(), w.shape[1]
        assert w.shape[0] == max(groups, 1) * N and len(taps) <= _lib.MAX_TAPS
        if groups > 1:
            assert group_rows * groups == M, (groups, group_rows, M)
            d.groups, d.group_rows, d.x_grouped = groups, group_rows, int(bool(x_grouped))
        transform = pre_scale is not None or x_is_f32
        nseg = 3 if (split if transform else x_lo_off) else 1
        assert w.shape[1] >= len(taps) * (2 if nseg == 3 else 1) * ((cin + 63) // 64) * 64, (w.shape, len(taps), nseg, cin)
        if transform:
            assert (pre_scale is None) == (pre_shift is None)
            assert x.dtype == (torch.float32 if x_is_f32 else torch.float16)
        d.M, d.N, d.taps = M, N, len(taps)
        for i, t in enumerate(taps):
            d.tap_off[i] = int(t)
            d.tap_coff[i] = int(tap_coff[i]) if tap_coff else 0
        d.s2d_c = int(s2d_c)
        oscale = getattr(w, "oscale", None)   # per-output-channel power-of-two pre-scaling of the packed weights (plan.pack_weight)
        if oscale is not None:
            assert oscale.dtype == torch.float32 and oscale.numel() == w.shape[0] and oscale.device == w.device
            d.out_scale = oscale.data_ptr()
        self._oscale_keep = oscale
        if head is not None:
            hw, hb, hs, ht, outs = head
            assert N == 64 and hw.dtype == torch.float32 and hw.shape == (max(groups, 1), 8, 64) and len(outs) == max(groups, 1)
            d.head_w, d.head_b, d.head_scale, d.head_shift = hw.data_ptr(), hb.data_ptr(), hs.data_ptr(), ht.data_ptr()
            for i, o in enumerate(outs):
                assert o.dtype == torch.float32 and o.is_contiguous() and o.shape[-1] <= 8
                d.head_out[i], d.head_nout[i], d.head_ld[i] = o.data_ptr(), o.shape[-1], o.shape[-1]
            self._head_keep = (hw, hb, hs, ht, outs)
        d.Hin, d.Win = grid
        d.y0, d.y1, d.x0, d.x1 = window
        d.Hd, d.Wd, d.dy, d.dx = dst_geom
        d.Hr, d.Wr, d.ry, d.rx = res_geom
        d.out16, d.ld16, d.coff16, d.lo16 = _ptr(out16), ld16, coff16, lo16
        d.out32, d.ld32, d.coff32 = _ptr(out32), ld32, coff32
        d.out16b, d.ld16b, d.coff16b, d.lo16b = _ptr(out16b), ld16b, coff16b, lo16b
        d.scale2, d.shift2, d.bias = _ptr(scale2), _ptr(shift2), _ptr(bias)
        d.res32, d.res16, d.ldres, d.coffres, d.reslo = _ptr(res32), _ptr(res16), ldres, coffres, reslo
        d.relu = int(bool(relu))
        d.pre_scale, d.pre_shift, d.pre_relu = _ptr(pre_scale), _ptr(pre_shift), int(bool(pre_relu))
        d.x_is_f32, d.split = int(bool(x_is_f32)), int(bool(split))
        self.desc = d
        self.dev = x.device
        self._keep = (x, w, out16, out32, out16b, scale2, shift2, bias, res32, res16, pre_scale, pre_shift)
        self.flops = 2.0 * M * N * len(taps) * cin
        self.mma_flops = self.flops * nseg
        # minimum HBM traffic of the launch (every operand read once, every output written once) — bench.py's HBM-bound / tensor-bound split
        in_b = float(M) * cin * (4 if x_is_f32 else (4 if nseg == 3 else 2))   # fp32 stream, or fp16 hi (+ lo) planes
        out_b = float(M) * N * ((4 if out32 is not None else 0) + (2 * (2 if lo16 else 1) if out16 is not None else 0)
                                + (2 * (2 if lo16b else 1) if out16b is not None else 0))
        res_b = float(M) * N * (4 if res32 is not None else ((4 if reslo else 2) if res16 is not None else 0))
        self.bytes = in_b + out_b + res_b + float(w.numel()) * 2

    def __call__(self):
        h = _lib.handle(self.dev.index or 0)
        _lib.check(h, _lib.lib().tiab_conv_gemm(h, C.byref(self.desc), _stream(self.dev)))


def im2col(src, src_is_u8, u8_scale, src_lo, n, h_, w_, c, ldsrc, kh, kw, stride, pad_t, pad_l, ho, wo, out, ldout, kpad, out_lo):
    h = _lib.handle(src.device.index or 0)
    _lib.check(h, _lib.lib().tiab_im2col(h, src.data_ptr(), int(src_is_u8), float(u8_scale), src_lo, n, h_, w_, c, ldsrc, kh, kw,
                                         stride, pad_t, pad_l, ho, wo, out.data_ptr(), ldout, kpad, out_lo, _stream(src.device)))


def bn_relu(src, src_is_f32, ldsrc, coffsrc, src_lo, scale, shift, relu, rows, c, dst, lddst, coffdst, dst_lo):
    h = _lib.handle(src.device.index or 0)
    _lib.check(h, _lib.lib().tiab_bn_relu(h, src.data_ptr(), int(src_is_f32), ldsrc, coffsrc, src_lo, _ptr(scale), _ptr(shift),
                                          int(relu), rows, c, dst.data_ptr(), lddst, coffdst, dst_lo, _stream(src.device)))


def upsample_add(a: Act, ha, wa, hag, wag, ay, ax, b: Act | None, hb, wb, by, bx, n, c, scale, shift, relu, dst: Act, hd, wd, dy, dx):
    h = _lib.handle(a.dev.index or 0)
    _lib.check(h, _lib.lib().tiab_upsample_add(
        h, a.t.data_ptr(), ha, wa, hag, wag, ay, ax, a.ld, a.lo, None if b is None else b.t.data_ptr(), hb, wb, by, bx,
        0 if b is None else b.ld, 0 if b is None else b.lo, n, c, _ptr(scale), _ptr(shift), int(relu), dst.t.data_ptr(), hd, wd,
        dy, dx, dst.ld, dst.lo, _stream(a.dev)))


def maxpool3x3s2(src: Act, n, h_, w_, c, dst: Act, hd, wd, dy, dx):
    h = _lib.handle(src.dev.index or 0)
    _lib.check(h, _lib.lib().tiab_maxpool3x3s2(h, src.t.data_ptr(), n, h_, w_, c, src.ld, src.lo, dst.t.data_ptr(), hd, wd, dy, dx,
                                               dst.ld, dst.lo, _stream(src.dev)))


def pointwise_head(x: Act, rows, c, scale, shift, relu, wt, bias, nout, out, ldout, coffout):
    h = _lib.handle(x.dev.index or 0)
    _lib.check(h, _lib.lib().tiab_pointwise_head(h, x.t.data_ptr(), rows, c, x.ld, x.lo, _ptr(scale), _ptr(shift), int(relu),
                                                 wt.data_ptr(), _ptr(bias), nout, out.data_ptr(), ldout, coffout, _stream(x.dev)))


def gap_fc_softmax(x: Act, n, hw, c, wt, bias, ncls, feat_out, probs):
    h = _lib.handle(x.dev.index or 0)
    _lib.check(h, _lib.lib().tiab_gap_fc_softmax(h, x.t.data_ptr(), n, hw, c, x.ld, x.lo, wt.data_ptr(), bias.data_ptr(), ncls,
                                                 _ptr(feat_out), probs.data_ptr(), _stream(x.dev)))


def hovernet_heads(npl, hvl, tpl, nt, rows, np_out, hv_out, tp_out):
    h = _lib.handle(npl.device.index or 0)
    _lib.check(h, _lib.lib().tiab_hovernet_heads(h, npl.data_ptr(), hvl.data_ptr(), _ptr(tpl), nt, rows, np_out.data_ptr(),
                                                 hv_out.data_ptr(), _ptr(tp_out), _stream(npl.device)))


def softmax_up2_crop(logits, n, h_, w_, c, out):
    h = _lib.handle(logits.device.index or 0)
    _lib.check(h, _lib.lib().tiab_softmax_up2_crop(h, logits.data_ptr(), n, h_, w_, c, out.data_ptr(), _stream(logits.device)))
