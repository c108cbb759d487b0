"""Device execution plans: a network forward as a fixed list of kernel launches over pre-allocated
NHWC buffers, replayed through a CUDA graph.

Design (DESIGN.md §conv engine): activations fp16, residual streams fp32, accumulation fp32 in TMEM.
BatchNorm that FOLLOWS a convolution is folded into the packed fp16 weights (+ fp32 bias in the
epilogue); BatchNorm+ReLU that PRECEDES a convolution (pre-activation blocks) is produced by the
epilogue of the layer that writes the tensor (second output `out16b`) or, when a tensor is consumed
under several different BNs (HoVer-Net dense blocks), by a small elementwise kernel.

Precision modes
  "fp16"    one tensor-core pass, fp16 operands (11-bit mantissa == TF32, what cuDNN uses by default
            for the reference's own CUDA path).  ~3e-4 relative error PER conv layer.
  "fp16x3"  every activation is carried as hi + lo fp16 planes and every weight as W_hi + W_lo; the
            GEMM runs A_hi*W_hi + A_lo*W_hi + A_hi*W_lo as one K-concatenated contraction (3x the MMA
            work, ~22-bit operands).  This is the mode that meets 1e-3 end-to-end through 60 layers.
"""
from __future__ import annotations

import torch

from tiatoolbox_b200.models import ops
from tiatoolbox_b200.models.architecture._params import bn_scale_shift
from tiatoolbox_b200.models.ops import Act


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _split_pack(packed64: torch.Tensor, split: bool) -> torch.Tensor:
    """[Cout, taps, Cpad] float64 -> fp16 [Cout, taps*Cpad] or, split, [Cout, taps*2*Cpad] = per tap [W_hi | W_lo]."""
    co, taps, cp = packed64.shape
    hi = packed64.to(torch.float16)
    if not split:
        return hi.reshape(co, taps * cp).contiguous()
    lo = (packed64 - hi.double()).to(torch.float16)
    return torch.stack([hi, lo], dim=2).reshape(co, taps * 2 * cp).contiguous()


def pack_weight(w: torch.Tensor, out_scale: torch.Tensor | None = None, groups: int = 1, split: bool = False, device=None,
                im2col: bool = False) -> torch.Tensor:
    """[Cout, Cin/groups, KH, KW] fp32 -> packed fp16 K-major weight matrix with the BN scale folded.

    Default layout: tap-major, each tap's channels zero padded to a multiple of 64 (shifted-GEMM path).
    im2col=True: one "tap" whose K = KH*KW*Cin (order kh, kw, c) padded to a multiple of 64.
    Grouped convolutions are expanded to a dense block-diagonal matrix (HoVer-Net's g=4 3x3 convs
    are 1 % of the FLOPs)."""
    w = w.detach().double().cpu()
    co, cig, kh, kw = w.shape
    cin = cig * groups
    if groups > 1:
        full = torch.zeros(co, cin, kh, kw, dtype=torch.float64)
        cog = co // groups
        for g in range(groups):
            full[g * cog:(g + 1) * cog, g * cig:(g + 1) * cig] = w[g * cog:(g + 1) * cog]
        w = full
    if out_scale is not None:
        w = w * out_scale.double().cpu().view(-1, 1, 1, 1)
    if im2col:
        k = kh * kw * cin
        packed = torch.zeros(co, 1, ceil_to(k, 64), dtype=torch.float64)
        packed[:, 0, :k] = w.permute(0, 2, 3, 1).reshape(co, k)
    else:
        cp = ceil_to(cin, 64)
        packed = torch.zeros(co, kh * kw, cp, dtype=torch.float64)
        packed[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, cin)
    # per-output-channel power-of-two pre-scaling (undone exactly by the epilogue's out_scale): keeps W_lo = W - fp16(W) inside
    # fp16's normal range for small-magnitude weights, which would otherwise leave the split product with ~14 instead of ~22 bits
    rowmax = packed.abs().amax(dim=(1, 2)).clamp_min(1e-30)
    pre = torch.exp2(torch.floor(torch.log2(64.0 / rowmax))).clamp(2.0 ** -16, 2.0 ** 24)
    out = _split_pack(packed * pre.view(-1, 1, 1), split).to(device)
    out.oscale = (1.0 / pre).to(torch.float32).to(device)
    return out


class Plan:
    """A list of launches + the buffers they touch."""

    def __init__(self, device, batch: int, precision: str = "fp16x3"):
        if precision not in ("fp16", "fp16x3"):
            msg = f"unknown precision {precision!r}"
            raise ValueError(msg)
        self.dev = torch.device(device)
        self.B = batch
        self.precision = precision
        self.split = precision == "fp16x3"
        self.ops: list = []
        self.bufs: dict[str, torch.Tensor] = {}
        self.flops = 0.0       # useful (algorithmic) FLOPs of one forward
        self.mma_flops = 0.0   # FLOPs actually issued to the tensor cores (incl. split segments, discarded rows)
        self.graph = None
        self._keep: list = []

    # ---- memory ----
    def buf(self, name: str, shape, dtype=torch.float16) -> torch.Tensor:
        t = torch.zeros(tuple(shape), dtype=dtype, device=self.dev)
        self.bufs[name] = t
        return t

    def act(self, name: str, lead_shape, C: int) -> Act:
        """fp16 activation buffer [*lead_shape, planes*C] (hi plane, then lo plane in split mode)."""
        planes = 2 if self.split else 1
        t = self.buf(name, (*lead_shape, planes * C))
        return Act(t, C, planes * C, C if self.split else 0)

    def const(self, t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        t = t.detach().to(dtype).contiguous().to(self.dev)
        self._keep.append(t)
        return t

    def weight(self, w, out_scale=None, groups=1, im2col=False) -> torch.Tensor:
        return pack_weight(w, out_scale, groups, self.split, self.dev, im2col)

    def fold_bn(self, sd, conv_name: str, bn_name: str, groups: int = 1, im2col: bool = False):
        """conv followed by BN: packed weight with the BN scale folded + fp32 bias (= BN shift)."""
        scale, shift = bn_scale_shift(sd, bn_name)
        return self.weight(sd[conv_name + ".weight"], scale, groups, im2col), self.const(shift)

    def bytes_allocated(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())

    # ---- launches ----
    def add(self, fn) -> None:
        self.ops.append(fn)

    @staticmethod
    def _out_kw(kw: dict) -> dict:
        """Expand Act-valued out16 / out16b / res16 keyword arguments into raw descriptor fields."""
        for key, ld, lo in (("out16", "ld16", "lo16"), ("out16b", "ld16b", "lo16b"), ("res16", "ldres", "reslo")):
            a = kw.get(key)
            if isinstance(a, Act):
                kw[key] = a.t
                kw[ld] = a.ld
                kw[lo] = a.lo
        return kw

    def _x_kw(self, x, pre) -> dict:
        """Operand-A description: an Act read by TMA, or (with `pre` = (scale, shift, relu) and/or an fp32 tensor) the
        transform path that applies the pre-activation BN+ReLU while staging the operand."""
        if isinstance(x, Act):
            kw = {"x": x.t, "x_ld": x.ld, "x_lo_off": x.lo}
        else:  # raw fp32 [.., C] tensor (residual stream)
            assert x.dtype == torch.float32
            kw = {"x": x, "x_ld": x.shape[-1], "x_lo_off": 0, "x_is_f32": True}
        if pre is not None or not isinstance(x, Act):
            kw["split"] = self.split
            if pre is not None:
                kw["pre_scale"], kw["pre_shift"], kw["pre_relu"] = pre
        return kw

    def conv(self, *, x, grid, cin: int, w: torch.Tensor, k: int, centered: bool, N: int, window, dst_geom=None, pre=None, nimg=None,
             groups: int = 1, x_grouped: bool = True, **kw) -> ops.ConvOp:
        """Shifted-GEMM conv over the (H, W) grid of the NHWC activation `x` (nimg images, default the plan's batch).
        groups > 1: the images are `groups` equal blocks, block g convolved with weight rows [g*N, (g+1)*N) of `w`
        (x_grouped=False: every block reads the same nimg/groups input images)."""
        H, W = grid
        nimg = nimg or self.B
        M = nimg * H * W
        if centered:
            taps = [(kh - k // 2) * W + (kw_ - k // 2) for kh in range(k) for kw_ in range(k)]
        else:
            taps = [kh * W + kw_ for kh in range(k) for kw_ in range(k)]
        gkw = {"groups": groups, "group_rows": M // groups, "x_grouped": x_grouped} if groups > 1 else {}
        op = ops.ConvOp(x_rows=M if x_grouped else M // groups, cin=cin, w=w, taps=taps, M=M, N=N, grid=(H, W), window=window,
                        dst_geom=dst_geom or (H, W, 0, 0), **gkw, **self._x_kw(x, pre), **self._out_kw(kw))
        y0, y1, x0, x1 = window
        self.flops += 2.0 * nimg * (y1 - y0) * (x1 - x0) * N * (k * k) * cin
        self.mma_flops += op.mma_flops
        self.ops.append(op)
        return op

    def conv_s2d(self, *, x: Act, grid, cin: int, w: torch.Tensor, k: int, N: int, out_hw, **kw) -> ops.ConvOp:
        """Stride-2 kxk convolution without im2col: `x` is the space-to-depth form of the (padded) input (cells grid = (Hs, Ws),
        4*cin channels per plane, written by the producing conv with `s2d_c=cin`), so input pixel (2*yo + kh, 2*xo + kw) is
        cell (yo + kh//2, xo + kw//2), channel block (kh%2)*2 + kw%2 -- a stride-1 shifted GEMM whose taps carry a channel offset."""
        Hs, Ws = grid
        Ho, Wo = out_hw
        assert x.C == 4 * cin and Hs >= Ho + (k - 1) // 2 and Ws >= Wo + (k - 1) // 2
        taps = [(kh // 2) * Ws + (kw_ // 2) for kh in range(k) for kw_ in range(k)]
        coff = [((kh % 2) * 2 + (kw_ % 2)) * cin for kh in range(k) for kw_ in range(k)]
        M = self.B * Hs * Ws
        op = ops.ConvOp(x=x.t, x_rows=M, x_ld=x.ld, x_lo_off=x.lo, cin=cin, w=w, taps=taps, tap_coff=coff, M=M, N=N, grid=(Hs, Ws),
                        window=(0, Ho, 0, Wo), dst_geom=(Ho, Wo, 0, 0), **self._out_kw(kw))
        self.flops += 2.0 * self.B * Ho * Wo * N * (k * k) * cin
        self.mma_flops += op.mma_flops
        self.ops.append(op)
        return op

    def weight_groups(self, ws, **kw) -> torch.Tensor:
        """Packed weights of several same-shape convolutions stacked along the output-channel axis (grouped launches)."""
        parts = [self.weight(w, **kw) for w in ws]
        out = torch.cat(parts, dim=0).contiguous()
        out.oscale = torch.cat([p.oscale for p in parts]).contiguous()
        return out

    def fold_bn_groups(self, sd, conv_names, bn_names):
        packed, biases = zip(*(self.fold_bn(sd, c, b) for c, b in zip(conv_names, bn_names)))
        out = torch.cat(packed, dim=0).contiguous()
        out.oscale = torch.cat([p.oscale for p in packed]).contiguous()
        return out, self.const(torch.cat([b.cpu() for b in biases]))

    def gemm(self, *, x, rows: int, cin: int, w: torch.Tensor, N: int, pre=None, taps=(0,), **kw) -> ops.ConvOp:
        """Plain [rows, cin] x [N, cin]^T GEMM (1x1 conv on a dense matrix / im2col output)."""
        op = ops.ConvOp(x_rows=rows, cin=cin, w=w, taps=list(taps), M=rows, N=N, grid=(1, rows),
                        window=(0, 1, 0, rows), dst_geom=kw.pop("dst_geom", (1, rows, 0, 0)), **self._x_kw(x, pre), **self._out_kw(kw))
        self.flops += 2.0 * rows * N * cin
        self.mma_flops += op.mma_flops
        self.ops.append(op)
        return op

    def im2col(self, src, src_is_u8, u8_scale, src_lo, n, h, w, c, ldsrc, k, stride, pad, ho, out: Act, kpad):
        self.add(lambda: ops.im2col(src, src_is_u8, u8_scale, src_lo, n, h, w, c, ldsrc, k, k, stride, pad, pad, ho, ho, out.t, out.ld,
                                    kpad, out.lo))

    def stem_u8(self, sd, conv_name: str, bn_name: str, *, inp: torch.Tensor, u8_scale: float, hw, k: int, stride: int, pad: int,
                out_hw, N: int, out16: Act, name: str = "stem.im2col") -> None:
        """First convolution straight from the uint8 NHWC batch: im2col (K = k*k*3 padded to 64) + GEMM + BN + ReLU.

        The input scale (1/255) is split into a power of two, applied to the bytes (byte * 2^-8 is exact in fp16, so the
        gathered operand needs no lo plane), and a factor in [0.7, 1.4] folded into the weights (folding all of 1/255 there
        would push W_lo = W - fp16(W) below fp16's subnormal range and cost ~8 bits).  Split precision then only has to cover
        the weights -- A*W_hi + A*W_lo, run as two zero-offset "taps" over the same rows (2 instead of 3 MMA segments, half
        the im2col bytes)."""
        import math

        H, W = hw
        Ho, Wo = out_hw
        scale, shift = bn_scale_shift(sd, bn_name)
        p2 = 2.0 ** round(math.log2(float(u8_scale)))
        w = self.weight(sd[conv_name + ".weight"], scale.double() * (float(u8_scale) / p2), im2col=True)
        kpad = ceil_to(k * k * 3, 64)
        rows = self.B * Ho * Wo
        gathered = Act(self.buf(name, (rows, kpad)), kpad, kpad, 0)
        self.add(lambda: ops.im2col(inp, True, p2, 0, self.B, H, W, 3, 3, k, k, stride, pad, pad, Ho, Wo, gathered.t, kpad, kpad, 0))
        self.gemm(x=gathered, rows=rows, cin=kpad, w=w, N=N, bias=self.const(shift), relu=True, out16=out16,
                  taps=(0, 0) if self.split else (0,))
        self.flops -= 2.0 * rows * N * (kpad - k * k * 3)   # count the real K, not the padding

    def bn_relu(self, src: Act, c: int, scale, shift, rows: int, dst: Act, relu=True):
        self.add(lambda: ops.bn_relu(src.t, False, src.ld, 0, src.lo, scale, shift, relu, rows, c, dst.t, dst.ld, 0, dst.lo))

    # ---- execution ----
    def run_eager(self) -> None:
        for op in self.ops:
            op()

    def capture(self) -> None:
        """Capture the whole launch list into one CUDA graph (removes ~150 launch latencies per step)."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.run_eager()  # warm-up (function attributes, tensor-map driver entry point)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_eager()
        self.graph = g

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_eager()

    @property
    def num_launches(self) -> int:
        return len(self.ops)
