"""Thin torch-tensor wrappers over the convolution-engine entry points of libtiab200.so.

Everything here is plumbing (pointers, pitches, geometry); the arithmetic is in csrc/conv.cu.
Tensors are NHWC and live on one CUDA device; launches go to torch's current stream so that the
whole network can be captured in a CUDA graph.
"""
from __future__ import annotations

import ctypes as C

import torch

from tiatoolbox_b200 import _lib


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


class ConvOp:
    """One shifted-GEMM convolution launch (see tiab_conv_desc_t)."""

    def __init__(self, *, x: torch.Tensor, x_rows: int, x_ld: int, cin: int, w: torch.Tensor, taps: list[int],
                 M: int, N: int, grid: tuple[int, int], window: tuple[int, int, int, int],
                 dst_geom: tuple[int, int, int, int], out16=None, ld16=0, coff16=0, out32=None, ld32=0, coff32=0,
                 out16b=None, ld16b=0, coff16b=0, scale2=None, shift2=None, bias=None, res32=None, res16=None,
                 ldres=0, coffres=0, res_geom=(1, 1, 0, 0), relu=False):
        d = _lib.ConvDesc()
        d.x, d.x_rows, d.x_ld, d.cin = x.data_ptr(), x_rows, x_ld, cin
        assert w.dtype == torch.float16 and w.dim() == 2 and w.is_contiguous()
        d.w, d.w_ld = w.data_ptr(), w.shape[1]
        assert w.shape[0] == N and len(taps) <= _lib.MAX_TAPS
        d.M, d.N, d.taps = M, N, len(taps)
        for i, t in enumerate(taps):
            d.tap_off[i] = int(t)
        d.Hin, d.Win = grid
        d.y0, d.y1, d.x0, d.x1 = window
        d.Hd, d.Wd, d.dy, d.dx = dst_geom
        d.Hr, d.Wr, d.ry, d.rx = res_geom
        d.out16, d.ld16, d.coff16 = _ptr(out16), ld16, coff16
        d.out32, d.ld32, d.coff32 = _ptr(out32), ld32, coff32
        d.out16b, d.ld16b, d.coff16b = _ptr(out16b), ld16b, coff16b
        d.scale2, d.shift2, d.bias = _ptr(scale2), _ptr(shift2), _ptr(bias)
        d.res32, d.res16, d.ldres, d.coffres = _ptr(res32), _ptr(res16), ldres, coffres
        d.relu = int(bool(relu))
        self.desc = d
        self.dev = x.device
        self._keep = (x, w, out16, out32, out16b, scale2, shift2, bias, res32, res16)
        self.flops = 2.0 * M * N * len(taps) * cin

    def __call__(self):
        h = _lib.handle(self.dev.index or 0)
        _lib.check(h, _lib.lib().tiab_conv_gemm(h, C.byref(self.desc), _stream(self.dev)))


def im2col(src, src_is_u8, u8_scale, n, h_, w_, c, ldsrc, kh, kw, stride, pad_t, pad_l, ho, wo, out, kpad):
    h = _lib.handle(src.device.index or 0)
    _lib.check(h, _lib.lib().tiab_im2col(h, src.data_ptr(), int(src_is_u8), float(u8_scale), n, h_, w_, c, ldsrc, kh, kw,
                                         stride, pad_t, pad_l, ho, wo, out.data_ptr(), kpad, _stream(src.device)))


def bn_relu(src, src_is_f32, ldsrc, coffsrc, scale, shift, relu, rows, c, dst, lddst, coffdst):
    h = _lib.handle(src.device.index or 0)
    _lib.check(h, _lib.lib().tiab_bn_relu(h, src.data_ptr(), int(src_is_f32), ldsrc, coffsrc, _ptr(scale), _ptr(shift),
                                          int(relu), rows, c, dst.data_ptr(), lddst, coffdst, _stream(src.device)))


def upsample_add(a, ha, wa, hag, wag, ay, ax, b, hb, wb, by, bx, n, c, scale, shift, relu, dst, hd, wd, dy, dx, dst32=None):
    h = _lib.handle(a.device.index or 0)
    _lib.check(h, _lib.lib().tiab_upsample_add(h, a.data_ptr(), ha, wa, hag, wag, ay, ax, _ptr(b), hb, wb, by, bx, n, c, _ptr(scale),
                                               _ptr(shift), int(relu), dst.data_ptr(), hd, wd, dy, dx, _ptr(dst32),
                                               _stream(a.device)))


def maxpool3x3s2(src, n, h_, w_, c, dst):
    h = _lib.handle(src.device.index or 0)
    _lib.check(h, _lib.lib().tiab_maxpool3x3s2(h, src.data_ptr(), n, h_, w_, c, dst.data_ptr(), _stream(src.device)))


def pointwise_head(x, rows, c, ldx, scale, shift, relu, wt, bias, nout, out, ldout, coffout):
    h = _lib.handle(x.device.index or 0)
    _lib.check(h, _lib.lib().tiab_pointwise_head(h, x.data_ptr(), rows, c, ldx, _ptr(scale), _ptr(shift), int(relu),
                                                 wt.data_ptr(), _ptr(bias), nout, out.data_ptr(), ldout, coffout,
                                                 _stream(x.device)))


def gap_fc_softmax(x, n, hw, c, wt, bias, ncls, feat_out, probs):
    h = _lib.handle(x.device.index or 0)
    _lib.check(h, _lib.lib().tiab_gap_fc_softmax(h, x.data_ptr(), n, hw, c, wt.data_ptr(), bias.data_ptr(), ncls,
                                                 _ptr(feat_out), probs.data_ptr(), _stream(x.device)))


def u8_to_f16(src, n, h_, w_, c, scale, dst, hd, wd, dy, dx, cd):
    h = _lib.handle(src.device.index or 0)
    _lib.check(h, _lib.lib().tiab_u8_to_f16(h, src.data_ptr(), n, h_, w_, c, float(scale), dst.data_ptr(), hd, wd, dy, dx, cd,
                                            _stream(src.device)))


def hovernet_heads(npl, hvl, tpl, nt, rows, np_out, hv_out, tp_out):
    h = _lib.handle(npl.device.index or 0)
    _lib.check(h, _lib.lib().tiab_hovernet_heads(h, npl.data_ptr(), hvl.data_ptr(), _ptr(tpl), nt, rows, np_out.data_ptr(),
                                                 hv_out.data_ptr(), _ptr(tp_out), _stream(npl.device)))
