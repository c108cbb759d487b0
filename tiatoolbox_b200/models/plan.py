"""Device execution plans: a network forward as a fixed list of kernel launches over pre-allocated
NHWC buffers, replayed through a CUDA graph.

Design (DESIGN.md §conv engine): activations fp16, residual streams fp32, accumulation fp32 in TMEM.
BatchNorm that FOLLOWS a convolution is folded into the packed fp16 weights (+ fp32 bias in the
epilogue); BatchNorm+ReLU that PRECEDES a convolution (pre-activation blocks) is produced by the
epilogue of the layer that writes the tensor (second output `out16b`) or, when a tensor is consumed
under several different BNs (HoVer-Net dense blocks), by a small elementwise kernel.
"""
from __future__ import annotations

import torch

from tiatoolbox_b200.models import ops
from tiatoolbox_b200.models.architecture._params import bn_scale_shift


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pack_weight(w: torch.Tensor, out_scale: torch.Tensor | None = None, groups: int = 1, extra_k: torch.Tensor | None = None,
                device=None) -> torch.Tensor:
    """[Cout, Cin/groups, KH, KW] fp32 -> [Cout, taps * ceil64(Cin)] fp16, tap-major, BN scale folded.

    Grouped convolutions are expanded to a dense block-diagonal matrix (HoVer-Net's g=4 3x3 convs
    are 1 % of the FLOPs)."""
    w = w.detach().double()
    co, cig, kh, kw = w.shape
    cin = cig * groups
    if groups > 1:
        full = torch.zeros(co, cin, kh, kw, dtype=torch.float64)
        cog = co // groups
        for g in range(groups):
            full[g * cog:(g + 1) * cog, g * cig:(g + 1) * cig] = w[g * cog:(g + 1) * cog]
        w = full
    if out_scale is not None:
        w = w * out_scale.double().view(-1, 1, 1, 1)
    cp = ceil_to(cin, 64)
    packed = torch.zeros(co, kh * kw, cp, dtype=torch.float64)
    packed[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, cin)
    packed = packed.reshape(co, kh * kw * cp)
    return packed.to(torch.float16).contiguous().to(device)


class Plan:
    """A list of launches + the buffers they touch."""

    def __init__(self, device, batch: int):
        self.dev = torch.device(device)
        self.B = batch
        self.ops: list = []
        self.bufs: dict[str, torch.Tensor] = {}
        self.flops = 0.0
        self.graph = None
        self._keep: list = []

    # ---- memory ----
    def buf(self, name: str, shape, dtype=torch.float16) -> torch.Tensor:
        t = torch.zeros(tuple(shape), dtype=dtype, device=self.dev)
        self.bufs[name] = t
        return t

    def const(self, t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        t = t.detach().to(dtype).contiguous().to(self.dev)
        self._keep.append(t)
        return t

    def bytes_allocated(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())

    # ---- launches ----
    def add(self, fn) -> None:
        self.ops.append(fn)

    def conv(self, *, x: torch.Tensor, grid, cin: int, w: torch.Tensor, k: int, centered: bool, N: int, window,
             x_ld: int | None = None, x_rows: int | None = None, dst_geom=None, **kw) -> ops.ConvOp:
        """Shifted-GEMM conv over the (H, W) grid of the NHWC buffer `x` (batch = self.B)."""
        H, W = grid
        M = self.B * H * W
        if centered:
            taps = [(kh - k // 2) * W + (kw_ - k // 2) for kh in range(k) for kw_ in range(k)]
        else:
            taps = [kh * W + kw_ for kh in range(k) for kw_ in range(k)]
        op = ops.ConvOp(x=x, x_rows=x_rows if x_rows is not None else M, x_ld=x_ld if x_ld is not None else x.shape[-1], cin=cin,
                        w=w, taps=taps, M=M, N=N, grid=(H, W), window=window, dst_geom=dst_geom or (H, W, 0, 0), **kw)
        # useful FLOPs: only the valid output window
        y0, y1, x0, x1 = window
        self.flops += 2.0 * self.B * (y1 - y0) * (x1 - x0) * N * (k * k) * cin
        self.ops.append(op)
        return op

    def gemm(self, *, x: torch.Tensor, rows: int, cin: int, w: torch.Tensor, N: int, x_ld: int | None = None, **kw) -> ops.ConvOp:
        """Plain [rows, cin] x [N, cin]^T GEMM (1x1 conv on a dense matrix / im2col output)."""
        op = ops.ConvOp(x=x, x_rows=rows, x_ld=x_ld if x_ld is not None else x.shape[-1], cin=cin, w=w, taps=[0], M=rows, N=N,
                        grid=(1, rows), window=(0, 1, 0, rows), dst_geom=kw.pop("dst_geom", (1, rows, 0, 0)), **kw)
        self.flops += 2.0 * rows * N * cin
        self.ops.append(op)
        return op

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


def fold_bn(sd, conv_name: str, bn_name: str, groups: int = 1, device=None):
    """conv followed by BN: packed fp16 weight with the BN scale folded + fp32 bias (= BN shift)."""
    scale, shift = bn_scale_shift(sd, bn_name)
    w = pack_weight(sd[conv_name + ".weight"], scale, groups, device=device)
    return w, shift.to(device)
