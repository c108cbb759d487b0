"""Layer-by-layer comparison of the HoVer-Net device plan with the fp32 torch reference (debug aid,
run on the GPU box:  PYTHONPATH=. python tools/debug_hovernet.py [fp16x3|fp16] [fast|original])."""
import sys

import numpy as np
import torch

from tests.cnn_common import hovernet_test_weights, rel_l2
from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
from oracle.torch_reference import hovernet_forward_reference
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec
from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan
from tiatoolbox_b200.synthetic import he_tiles

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
S = 256 if mode == "fast" else 270
B = 2
x = torch.from_numpy(he_tiles(B, S, S, seed=0)).cuda()
xf = x.permute(0, 3, 1, 2).float()
if mode == "fast":
    sd = hovernet_test_weights("fast", 6)
else:
    sdg = {k: v.cuda() for k, v in seeded_state_dict(hovernet_spec(3, 6, mode), seed=0).items()}
    sdg.update(calibrate(sdg, lambda s, inp: hovernet_forward_reference(s, inp, mode, 6), xf))
    sd = {k: v.cpu() for k, v in sdg.items()}
plan = HoVerNetPlan(sd, batch=B, device="cuda:0", mode=mode, num_types=6, precision=prec)
print("precision", prec, "mode", mode)
plan.forward(x)
torch.cuda.synchronize()
cap = {}
with torch.inference_mode():
    ref = hovernet_forward_reference({k: v.cuda() for k, v in sd.items()}, xf, mode, 6, cap)


def cmp(name, got, r, win=None):
    r = r.permute(0, 2, 3, 1).float().cpu().numpy()
    got = got.float().cpu().numpy()
    if win is not None:
        lo, n = win
        got = got[:, lo:lo + n, lo:lo + n]
    print(f"{name:14s} shape {got.shape} rel_l2 {rel_l2(got, r):.3e}  max|ref| {np.abs(r).max():.3f}  finite {np.isfinite(got).all()}")


def val(name, c, imgs=None):
    t = plan.bufs[name]
    if imgs is not None:
        t = t[imgs[0]:imgs[1]]
    planes = 2 if plan.split else 1
    C = t.shape[-1] // planes
    v = t[..., :C].float()
    if plan.split:
        v = v + t[..., C:].float()
    return v[..., :c]


_cat = plan.bufs["d0.cat"]          # [t2 | x0] per plane; conv0's output is the second 64-channel slice
_C = _cat.shape[-1] // (2 if plan.split else 1)
cmp("x0", _cat[..., 64:128].float() + (_cat[..., _C + 64:_C + 128].float() if plan.split else 0), cap["x0"])
for n in ("d0", "d1", "d2", "d3"):
    cmp(n, val(f"{n}.out", 10**6), cap[n])
cmp("dbot", val("dbot", 1024), cap["dbot"])
cmp("u3in", val("u3in", 1024), cap["u3in"])
# decoder buffers hold the branches as consecutive image blocks (tp, np, hv): hv is block 2
hv = (2 * B, 3 * B)
k = 5 if mode == "original" else 3
h = k // 2
G3 = plan.bufs["cb3"].shape[1]
G2 = plan.bufs["cb2"].shape[1]
cmp("hv.u3.conva", val("cb3", 256, hv), cap["hv.u3.conva"], (h, G3 - 2 * h))
cmp("hv.u3out", val("u3out", 512, hv), cap["hv.u3out"], (9 * h, G3 - 18 * h))
cmp("hv.u2in", val("u2in", 512, hv), cap["hv.u2in"])
cmp("hv.u2out", val("u2out", 256, hv), cap["hv.u2out"], (5 * h, G2 - 10 * h))
G1 = plan.bufs["u1in_padded"].shape[1]
cmp("hv.u1in", val("u1in_padded", 256, hv), cap["hv.u1in"], ((G1 - plan.So) // 2, plan.So))
for b in ("tp", "np", "hv"):
    cmp(f"logits {b}", plan.logits[b], ref[b])

# ---- the first decoder conv in isolation: reference input (hi/lo split), hv weights, plain and grouped launches ----
import torch.nn.functional as F  # noqa: E402

from tiatoolbox_b200.models import ops  # noqa: E402
from tiatoolbox_b200.models.plan import pack_weight  # noqa: E402

u3 = cap["u3in"].permute(0, 2, 3, 1).contiguous()          # (B, G3, G3, 1024) fp32
hi = u3.half()
lo = (u3 - hi.float()).half()
xa = torch.cat([hi, lo], dim=-1).contiguous()
names = ["tp", "np", "hv"]
ws = [sd[f"decoder.{n}.u3.conva.weight"].cuda() for n in names]
M1 = B * G3 * G3
taps = [(kh - h) * G3 + (kw - h) for kh in range(k) for kw in range(k)]
for label, wl in (("plain hv", [ws[2]]), ("grouped", ws)):
    R = len(wl)
    wp = torch.cat([pack_weight(w, split=True, device="cuda") for w in wl], dim=0).contiguous()
    out = torch.zeros(R * B, G3, G3, 256, device="cuda")
    kw = dict(groups=R, group_rows=M1, x_grouped=False) if R > 1 else {}
    op = ops.ConvOp(x=xa, x_rows=M1, x_ld=2048, cin=1024, x_lo_off=1024, w=wp, taps=taps, M=R * M1, N=256, grid=(G3, G3),
                    window=(h, G3 - h, h, G3 - h), dst_geom=(G3, G3, 0, 0), out32=out, ld32=256, **kw)
    op()
    torch.cuda.synchronize()
    for r, w in enumerate(wl):
        ref64 = F.conv2d((hi.float() + lo.float()).double().permute(0, 3, 1, 2), w.double()).permute(0, 2, 3, 1).float()
        got = out[r * B:(r + 1) * B, h:G3 - h, h:G3 - h]
        print(f"isolated conva {label} group {r}: rel_l2 vs fp64 {((got - ref64).norm() / ref64.norm()).item():.3e}   "
              f"fp32 torch vs fp64 {((F.conv2d(u3.permute(0, 3, 1, 2), w).permute(0, 2, 3, 1) - ref64).norm() / ref64.norm()).item():.3e}")

# same isolated launch, but fed with the PLAN's own u3in planes; then the plan's cb3 blocks against both
pu = plan.bufs["u3in"]
outp = torch.zeros(3 * B, G3, G3, 256, device="cuda")
wp = torch.cat([pack_weight(w, split=True, device="cuda") for w in ws], dim=0).contiguous()
ops.ConvOp(x=pu, x_rows=M1, x_ld=2048, cin=1024, x_lo_off=1024, w=wp, taps=taps, M=3 * M1, N=256, grid=(G3, G3),
           window=(h, G3 - h, h, G3 - h), dst_geom=(G3, G3, 0, 0), out32=outp, ld32=256, groups=3, group_rows=M1, x_grouped=False)()
torch.cuda.synchronize()
for r, n in enumerate(names):
    refc = cap[f"{n}.u3.conva"].permute(0, 2, 3, 1)
    iso = outp[r * B:(r + 1) * B, h:G3 - h, h:G3 - h]
    pl = val("cb3", 256, (r * B, (r + 1) * B))[:, h:G3 - h, h:G3 - h]
    e_iso = ((iso - refc).norm() / refc.norm()).item()
    e_pl = ((pl - refc).norm() / refc.norm()).item()
    d = (pl - refc).abs()
    idx = torch.nonzero(d == d.max())[0].tolist()
    print(f"{n}: isolated(plan input) vs ref {e_iso:.3e}   plan cb3 vs ref {e_pl:.3e}   plan vs isolated {((pl - iso).norm() / iso.norm()).item():.3e}"
          f"   max|d| {d.max().item():.3e} at {idx}")
    per_ch = ((pl - refc) ** 2).sum(dim=(0, 1, 2)).sqrt() / (refc ** 2).sum(dim=(0, 1, 2)).sqrt()
    print("   worst channels", torch.topk(per_ch, 5).indices.tolist(), [f"{v:.2e}" for v in torch.topk(per_ch, 5).values.tolist()])

# ---- how does the first decoder conv treat the plan's input perturbation? ----
pa = val("u3in", 1024)                                   # plan's u3in (hi + lo)
da = pa - u3
print(f"u3in: |a| rms {u3.pow(2).mean().sqrt().item():.3f} mean {u3.mean().item():.3f}  |da| rms {da.pow(2).mean().sqrt().item():.3e} "
      f"mean {da.mean().item():.3e} max {da.abs().max().item():.3e}")
wv = ws[2]
ca = F.conv2d(u3.permute(0, 3, 1, 2), wv)
cd = F.conv2d(da.permute(0, 3, 1, 2), wv)
print(f"gain on a: {(ca.norm() / u3.norm()).item():.4f}   gain on da: {(cd.norm() / da.norm()).item():.4f}   "
      f"weights: mean {wv.mean().item():.3e} std {wv.std().item():.3e}  sum/|w|1 {(wv.sum() / wv.abs().sum()).item():.3e}")
ch_mean = u3.mean(dim=(0, 1, 2))
print(f"per-channel mean of a: rms {ch_mean.pow(2).mean().sqrt().item():.3f};  a minus channel means rms {(u3 - ch_mean).pow(2).mean().sqrt().item():.3f}")
rel_by_row = (da.pow(2).sum(dim=(0, 2, 3)).sqrt() / u3.pow(2).sum(dim=(0, 2, 3)).sqrt())
print("rel err of u3in by grid row (first 6, middle, last 6):", [f"{v:.1e}" for v in rel_by_row[:6].tolist()], f"{rel_by_row[G3 // 2].item():.1e}",
      [f"{v:.1e}" for v in rel_by_row[-6:].tolist()])

ref64 = F.conv2d(u3.double().permute(0, 3, 1, 2).contiguous(), wv.double()).float()
print(f"reference's own fp32 conva (inside the forward) vs fp64 on the same input: {((cap['hv.u3.conva'] - ref64).norm() / ref64.norm()).item():.3e}")
