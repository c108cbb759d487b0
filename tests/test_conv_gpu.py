"""GPU tests of the tcgen05 shifted-GEMM convolution kernel against plain PyTorch fp32 references
computed from the same fp16-rounded operands (so the only difference is fp32 summation order)."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    from tiatoolbox_b200.models import ops as _ops

    return _ops


def _pack_w(w_oihw, cin_pad):
    """[Cout, Cin, KH, KW] fp32 -> [Cout, taps*cin_pad] fp16 (tap-major, channels zero padded)."""
    import torch

    co, ci, kh, kw = w_oihw.shape
    w = torch.zeros(co, kh * kw, cin_pad, dtype=torch.float16, device=w_oihw.device)
    w[:, :, :ci] = w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).half()
    return w.reshape(co, kh * kw * cin_pad).contiguous()


@pytest.mark.parametrize("m,k,n", [(1000, 192, 64), (128, 64, 32), (4096 + 77, 288, 128), (777, 1024, 256), (300, 64, 512)])
def test_gemm_1x1(ops, m, k, n):
    import torch

    torch.manual_seed(0)
    dev = torch.device("cuda")
    x = torch.randn(m, k, device=dev).half()
    w = (torch.randn(n, k, device=dev) / k**0.5)
    bias = torch.randn(n, device=dev)
    kp = (k + 63) // 64 * 64
    wp = _pack_w(w.reshape(n, k, 1, 1), kp)
    out32 = torch.full((m, n), float("nan"), device=dev)
    out16 = torch.zeros((m, n), dtype=torch.float16, device=dev)
    op = ops.ConvOp(x=x, x_rows=m, x_ld=k, cin=k, w=wp, taps=[0], M=m, N=n, grid=(1, m), window=(0, 1, 0, m),
                    dst_geom=(1, m, 0, 0), out32=out32, ld32=n, out16=out16, ld16=n, bias=bias)
    op()
    torch.cuda.synchronize()
    ref = x.float() @ w.half().float().t() + bias
    assert torch.isfinite(out32).all()
    err = (out32 - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    assert (out16.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 20), (128, 32, 17), (96, 128, 33)])
def test_conv3x3_valid_with_epilogues(ops, cin, cout, hw):
    import torch
    import torch.nn.functional as F

    torch.manual_seed(1)
    dev = torch.device("cuda")
    nimg, h, w_ = 3, hw, hw + 3
    x = torch.randn(nimg, h, w_, cin, device=dev).half()          # NHWC
    wt = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
    bias = torch.randn(cout, device=dev) * 0.1
    res = torch.randn(nimg, h - 2, w_ - 2, cout, device=dev)      # fp32 residual on the output grid
    scale2 = torch.rand(cout, device=dev) + 0.5
    shift2 = torch.randn(cout, device=dev) * 0.1
    cin_pad = (cin + 63) // 64 * 64
    wp = _pack_w(wt, cin_pad)
    ho, wo = h - 2, w_ - 2
    # destination: interior of a zero padded buffer (pad 1) with extra channels (channel offset 32)
    cd = cout + 64
    dst16 = torch.zeros(nimg, ho + 2, wo + 2, cd, dtype=torch.float16, device=dev)
    dst32 = torch.zeros(nimg, ho, wo, cout, device=dev)
    dst16b = torch.zeros(nimg, ho, wo, cout, dtype=torch.float16, device=dev)
    taps = [kh * w_ + kw for kh in range(3) for kw in range(3)]
    op = ops.ConvOp(x=x, x_rows=nimg * h * w_, x_ld=cin, cin=cin, w=wp, taps=taps, M=nimg * h * w_, N=cout,
                    grid=(h, w_), window=(0, ho, 0, wo), dst_geom=(ho + 2, wo + 2, 1, 1), out16=dst16, ld16=cd, coff16=32,
                    out32=dst32, ld32=cout, out16b=dst16b, ld16b=cout, scale2=scale2, shift2=shift2, bias=bias,
                    res32=res, ldres=cout, res_geom=(ho, wo, 0, 0), relu=False)
    # out32 uses a different destination geometry than out16 -> launch twice with two descriptors instead
    op.desc.out32 = None
    op.desc.out16b = None
    op()
    op2 = ops.ConvOp(x=x, x_rows=nimg * h * w_, x_ld=cin, cin=cin, w=wp, taps=taps, M=nimg * h * w_, N=cout,
                     grid=(h, w_), window=(0, ho, 0, wo), dst_geom=(ho, wo, 0, 0), out32=dst32, ld32=cout,
                     out16b=dst16b, ld16b=cout, scale2=scale2, shift2=shift2, bias=bias, res32=res, ldres=cout,
                     res_geom=(ho, wo, 0, 0), relu=False)
    op2()
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float(), bias).permute(0, 2, 3, 1) + res
    tol = 3e-3 * max(1.0, ref.abs().max().item())
    assert (dst32 - ref).abs().max().item() < tol
    assert (dst16[:, 1:-1, 1:-1, 32:32 + cout].float() - ref).abs().max().item() < 3e-2
    # untouched: the zero border and the other channels
    assert dst16[:, 0].abs().max().item() == 0 and dst16[:, :, 0].abs().max().item() == 0
    assert dst16[..., :32].abs().max().item() == 0 and dst16[..., 32 + cout:].abs().max().item() == 0
    ref_b = torch.relu(ref * scale2 + shift2)
    assert (dst16b.float() - ref_b).abs().max().item() < 3e-2


def test_conv_relu_fp16_residual_and_many_tiles(ops):
    import torch
    import torch.nn.functional as F

    torch.manual_seed(2)
    dev = torch.device("cuda")
    nimg, h, w_, cin, cout = 4, 66, 66, 256, 256      # > 148 tiles: persistent loop + TMEM double buffering
    x = torch.randn(nimg, h, w_, cin, device=dev).half()
    wt = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
    res16 = torch.randn(nimg, h - 2, w_ - 2, cout, device=dev).half()
    wp = _pack_w(wt, cin)
    ho, wo = h - 2, w_ - 2
    dst16 = torch.zeros(nimg, ho, wo, cout, dtype=torch.float16, device=dev)
    taps = [kh * w_ + kw for kh in range(3) for kw in range(3)]
    op = ops.ConvOp(x=x, x_rows=nimg * h * w_, x_ld=cin, cin=cin, w=wp, taps=taps, M=nimg * h * w_, N=cout, grid=(h, w_),
                    window=(0, ho, 0, wo), dst_geom=(ho, wo, 0, 0), out16=dst16, ld16=cout, res16=res16, ldres=cout,
                    res_geom=(ho, wo, 0, 0), relu=True)
    op()
    torch.cuda.synchronize()
    ref = torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float()).permute(0, 2, 3, 1) + res16.float())
    assert (dst16.float() - ref).abs().max().item() < 3e-2
    # timing sanity (not a benchmark): 2*M*N*K flops
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        op()
    start.record()
    for _ in range(10):
        op()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / 10
    print(f"conv3x3 256->256 @66x66x4: {ms*1e3:.1f} us, {op.flops/ms/1e9:.1f} TFLOP/s")


@pytest.mark.parametrize("wmag", [1.0, 1e-4])
@pytest.mark.parametrize("m,k,n", [(1000, 192, 64), (515, 288, 128)])
def test_split_precision_gemm(ops, m, k, n, wmag):
    """3-term split (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo): matches the full fp32 product to ~1e-6 relative,
    and writes hi/lo output planes whose sum carries the fp32 result.  wmag = 1e-4: weights whose fp16 residual would
    underflow without the per-channel power-of-two pre-scaling (undone by the epilogue's out_scale)."""
    import torch

    from tiatoolbox_b200.models.plan import pack_weight

    torch.manual_seed(3)
    dev = torch.device("cuda")
    xf = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k**0.5 * wmag
    x = torch.zeros(m, 2 * k, dtype=torch.float16, device=dev)
    x[:, :k] = xf.half()
    x[:, k:] = (xf - xf.half().float()).half()
    wp = pack_weight(w.reshape(n, k, 1, 1), split=True, device=dev)
    out32 = torch.zeros(m, n, device=dev)
    out16 = torch.zeros(m, 2 * n, dtype=torch.float16, device=dev)
    op = ops.ConvOp(x=x, x_rows=m, x_ld=2 * k, x_lo_off=k, cin=k, w=wp, taps=[0], M=m, N=n, grid=(1, m), window=(0, 1, 0, m),
                    dst_geom=(1, m, 0, 0), out32=out32, ld32=n, out16=out16, ld16=2 * n, lo16=n)
    op()
    torch.cuda.synchronize()
    ref = (xf.double() @ w.double().t()).float()
    scale = ref.abs().max().item()
    assert (out32 - ref).abs().max().item() < 2e-5 * scale
    if wmag == 1.0:   # (outputs of magnitude 1e-4 are below fp16's normal range themselves: only the fp32 output is meaningful there)
        recon = out16[:, :n].float() + out16[:, n:].float()
        assert (recon - ref).abs().max().item() < 2e-5 * scale


@pytest.mark.parametrize("src_f32", [True, False])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("m,k,n", [(1000, 288, 128), (4096 + 50, 64, 256), (300, 512, 32), (148 * 128 * 3 + 17, 256, 64), (40000, 1024, 256)])
def test_transform_operand_mode(ops, src_f32, split, m, k, n):
    """A = relu(x*scale + shift) built inside the GEMM (pre-activation BN+ReLU fused into the operand path)."""
    import torch

    from tiatoolbox_b200.models.plan import pack_weight

    torch.manual_seed(4)
    dev = torch.device("cuda")
    xf = torch.randn(m, k, device=dev)
    scale = torch.rand(k, device=dev) + 0.5
    shift = torch.randn(k, device=dev) * 0.3
    w = torch.randn(n, k, device=dev) / k**0.5
    bias = torch.randn(n, device=dev) * 0.1
    ld = k + 64  # pitch larger than cin: stale channels beyond cin must be ignored
    if src_f32:
        x = torch.full((m, ld), 7.0, device=dev)
        x[:, :k] = xf
        xin = xf
        kw = dict(x=x, x_ld=ld, x_lo_off=0, x_is_f32=True)
    else:
        x = torch.full((m, 2 * ld), 7.0, dtype=torch.float16, device=dev)
        x[:, :k] = xf.half()
        x[:, ld:ld + k] = (xf - xf.half().float()).half()
        xin = x[:, :k].float() + x[:, ld:ld + k].float()
        kw = dict(x=x, x_ld=2 * ld, x_lo_off=ld)
    wp = pack_weight(w.reshape(n, k, 1, 1), split=split, device=dev)
    out32 = torch.zeros(m, n, device=dev)
    op = ops.ConvOp(x_rows=m, cin=k, w=wp, taps=[0], M=m, N=n, grid=(1, m), window=(0, 1, 0, m), dst_geom=(1, m, 0, 0), out32=out32,
                    ld32=n, bias=bias, pre_scale=scale, pre_shift=shift, pre_relu=True, split=split, **kw)
    op()
    torch.cuda.synchronize()
    a = torch.relu(xin * scale + shift)
    if split:
        ref = (a.double() @ w.double().t()).float() + bias
        tol = 3e-5
    else:
        ref = a.half().float() @ w.half().float().t() + bias
        tol = 2e-3
    err = (out32 - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, err


@pytest.mark.parametrize("transform", [False, True])
@pytest.mark.parametrize("k", [1, 3])
def test_window_tile_skipping(ops, k, transform):
    """A y-window that excludes whole 128-row tiles (and image-straddling tiles): outputs inside the window match the
    reference, everything outside is left untouched."""
    import torch
    import torch.nn.functional as F

    if transform and k != 1:
        pytest.skip("operand-transform mode is used for 1x1 convolutions only")
    torch.manual_seed(5)
    dev = torch.device("cuda")
    nimg, H, W, cin, cout = 5, 40, 24, 64, 64
    y0, y1, x0, x1 = 13, 29, 2, 21
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    wp = _pack_w(w, cin)
    out32 = torch.full((nimg, H, W, cout), -7.0, device=dev)
    taps = [(kh - k // 2) * W + (kw - k // 2) for kh in range(k) for kw in range(k)]
    if transform:
        xs = torch.randn(nimg, H, W, cin, device=dev)
        scale, shift = torch.rand(cin, device=dev) + 0.5, torch.randn(cin, device=dev)
        a = torch.relu(xs * scale + shift).half().float()
        op = ops.ConvOp(x=xs, x_rows=nimg * H * W, x_ld=cin, cin=cin, w=wp, taps=taps, M=nimg * H * W, N=cout, grid=(H, W),
                        window=(y0, y1, x0, x1), dst_geom=(H, W, 0, 0), out32=out32, ld32=cout, pre_scale=scale, pre_shift=shift,
                        pre_relu=True, x_is_f32=True, split=False)
    else:
        xh = torch.randn(nimg, H, W, cin, device=dev).half()
        a = xh.float()
        op = ops.ConvOp(x=xh, x_rows=nimg * H * W, x_ld=cin, cin=cin, w=wp, taps=taps, M=nimg * H * W, N=cout, grid=(H, W),
                        window=(y0, y1, x0, x1), dst_geom=(H, W, 0, 0), out32=out32, ld32=cout)
    op()
    torch.cuda.synchronize()
    ref = F.conv2d(a.permute(0, 3, 1, 2), w.half().float(), padding=k // 2).permute(0, 2, 3, 1)
    inside = torch.zeros(nimg, H, W, dtype=torch.bool, device=dev)
    inside[:, y0:y1, x0:x1] = True
    assert (out32[~inside] == -7.0).all()
    err = (out32[inside] - ref[inside]).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("mode", ["shared_input_3x3", "grouped_input_3x3", "grouped_transform_1x1"])
def test_grouped_launch(ops, mode):
    """Three same-shape convolutions with different weights in one grid (HoVer-Net's decoder branches): group rows are not
    a multiple of the 128-row tile, so every group ends in a partial tile."""
    import torch
    import torch.nn.functional as F

    torch.manual_seed(6)
    dev = torch.device("cuda")
    R, nimg, H, W, cin, cout = 3, 2, 19, 21, 64, 64          # group_rows = 2*19*21 = 798
    k = 1 if mode == "grouped_transform_1x1" else 3
    shared = mode == "shared_input_3x3"
    w = torch.randn(R, cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    bias = torch.randn(R * cout, device=dev) * 0.1
    wp = torch.cat([_pack_w(w[r], cin) for r in range(R)], dim=0).contiguous()
    taps = [(kh - k // 2) * W + (kw - k // 2) for kh in range(k) for kw in range(k)]
    y0, y1, x0, x1 = k // 2, H - k // 2, k // 2, W - k // 2
    M = R * nimg * H * W
    out32 = torch.full((R * nimg, H, W, cout), -7.0, device=dev)
    if mode == "grouped_transform_1x1":
        xs = torch.randn(R * nimg, H, W, cin, device=dev)
        scale, shift = torch.rand(R * cin, device=dev) + 0.5, torch.randn(R * cin, device=dev)
        a = torch.relu(xs.view(R, nimg, H, W, cin) * scale.view(R, 1, 1, 1, cin) + shift.view(R, 1, 1, 1, cin)).half().float()
        op = ops.ConvOp(x=xs, x_rows=M, x_ld=cin, cin=cin, w=wp, taps=taps, M=M, N=cout, grid=(H, W), window=(y0, y1, x0, x1),
                        dst_geom=(H, W, 0, 0), out32=out32, ld32=cout, bias=bias, pre_scale=scale, pre_shift=shift, pre_relu=True,
                        x_is_f32=True, split=False, groups=R, group_rows=nimg * H * W)
    else:
        nin = nimg if shared else R * nimg
        xh = torch.randn(nin, H, W, cin, device=dev).half()
        a = (xh.float().unsqueeze(0).expand(R, -1, -1, -1, -1) if shared else xh.float().view(R, nimg, H, W, cin))
        op = ops.ConvOp(x=xh, x_rows=nin * H * W, x_ld=cin, cin=cin, w=wp, taps=taps, M=M, N=cout, grid=(H, W),
                        window=(y0, y1, x0, x1), dst_geom=(H, W, 0, 0), out32=out32, ld32=cout, bias=bias, groups=R,
                        group_rows=nimg * H * W, x_grouped=not shared)
    op()
    torch.cuda.synchronize()
    got = out32.view(R, nimg, H, W, cout)
    for r in range(R):
        ref = F.conv2d(a[r].permute(0, 3, 1, 2), w[r].half().float(), bias[r * cout:(r + 1) * cout], padding=k // 2).permute(0, 2, 3, 1)
        err = (got[r][:, y0:y1, x0:x1] - ref[:, y0:y1, x0:x1]).abs().max().item()
        assert err < 2e-3 * max(1.0, ref.abs().max().item()), (r, err)
    if k == 3:
        assert (got[:, :, 0] == -7.0).all() and (got[:, :, :, 0] == -7.0).all()


@pytest.mark.parametrize("hw", [12, 11])
def test_stride2_conv_via_space_to_depth(ops, hw):
    """1x1 producer storing space-to-depth + 3x3 stride-2 consumer with per-tap channel offsets == TF-'same' stride-2 conv."""
    import torch
    import torch.nn.functional as F

    torch.manual_seed(7)
    dev = torch.device("cuda")
    nimg, H, cin, c1, cout = 2, hw, 64, 64, 64
    Ho = (H + 1) // 2
    pad_total = max((Ho - 1) * 2 + 3 - H, 0)
    pb, pe = pad_total // 2, pad_total - pad_total // 2
    Hs = Ho + 1
    x = torch.randn(nimg, H, H, cin, device=dev).half()
    w1 = torch.randn(c1, cin, 1, 1, device=dev) / cin**0.5
    w2 = torch.randn(cout, c1, 3, 3, device=dev) / (9 * c1) ** 0.5
    t1s = torch.zeros(nimg, Hs, Hs, 4 * c1, dtype=torch.float16, device=dev)
    M1 = nimg * H * H
    op1 = ops.ConvOp(x=x, x_rows=M1, x_ld=cin, cin=cin, w=_pack_w(w1, cin), taps=[0], M=M1, N=c1, grid=(H, H), window=(0, H, 0, H),
                     dst_geom=(Hs, Hs, pb, pb), out16=t1s, ld16=4 * c1, s2d_c=c1, relu=True)
    op1()
    taps = [(kh // 2) * Hs + (kw // 2) for kh in range(3) for kw in range(3)]
    coff = [((kh % 2) * 2 + (kw % 2)) * c1 for kh in range(3) for kw in range(3)]
    M2 = nimg * Hs * Hs
    out32 = torch.full((nimg, Ho, Ho, cout), -3.0, device=dev)
    op2 = ops.ConvOp(x=t1s, x_rows=M2, x_ld=4 * c1, cin=c1, w=_pack_w(w2, c1), taps=taps, tap_coff=coff, M=M2, N=cout, grid=(Hs, Hs),
                     window=(0, Ho, 0, Ho), dst_geom=(Ho, Ho, 0, 0), out32=out32, ld32=cout)
    op2()
    torch.cuda.synchronize()
    t1 = torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w1.half().float())).half().float()
    ref = F.conv2d(F.pad(t1, (pb, pe, pb, pe)), w2.half().float(), stride=2).permute(0, 2, 3, 1)
    assert ref.shape == out32.shape
    err = (out32 - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_fused_pointwise_head(ops):
    """3x3 conv (N = 64, two groups) with BN + ReLU + 1x1 conv + bias fused into the epilogue; only the fp32 logits are stored."""
    import torch
    import torch.nn.functional as F

    torch.manual_seed(8)
    dev = torch.device("cuda")
    R, nimg, H, W, cin = 2, 2, 20, 22, 64
    nouts = [2, 6]
    xh = torch.randn(R * nimg, H, W, cin, device=dev).half()
    w = torch.randn(R, 64, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
    wp = torch.cat([_pack_w(w[r], cin) for r in range(R)], dim=0).contiguous()
    hw = torch.zeros(R, 8, 64, device=dev)
    hb = torch.zeros(R, 8, device=dev)
    for r, oc in enumerate(nouts):
        hw[r, :oc] = torch.randn(oc, 64, device=dev) / 8.0
        hb[r, :oc] = torch.randn(oc, device=dev)
    hs, ht = torch.rand(R, 64, device=dev) + 0.5, torch.randn(R, 64, device=dev) * 0.2
    Ho, Wo = H - 2, W - 2
    outs = [torch.full((nimg, Ho, Wo, oc), -5.0, device=dev) for oc in nouts]
    taps = [kh * W + kw for kh in range(3) for kw in range(3)]
    M = R * nimg * H * W
    op = ops.ConvOp(x=xh, x_rows=M, x_ld=cin, cin=cin, w=wp, taps=taps, M=M, N=64, grid=(H, W), window=(0, Ho, 0, Wo),
                    dst_geom=(Ho, Wo, 0, 0), groups=R, group_rows=nimg * H * W, head=(hw, hb, hs.contiguous(), ht.contiguous(), outs))
    op()
    torch.cuda.synchronize()
    for r, oc in enumerate(nouts):
        xr = xh[r * nimg:(r + 1) * nimg].float().permute(0, 3, 1, 2)
        feat = F.conv2d(xr, w[r].half().float())                                   # valid 3x3
        act = torch.relu(feat * hs[r].view(1, 64, 1, 1) + ht[r].view(1, 64, 1, 1))
        ref = (F.conv2d(act, hw[r, :oc].view(oc, 64, 1, 1)) + hb[r, :oc].view(1, oc, 1, 1)).permute(0, 2, 3, 1)
        err = (outs[r] - ref).abs().max().item()
        assert err < 2e-3 * max(1.0, ref.abs().max().item()), (r, err)
