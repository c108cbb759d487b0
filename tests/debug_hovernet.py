"""Layer-by-layer comparison of the HoVer-Net device plan with the fp32 torch reference (debug aid,
run on the GPU box:  python -m tests.debug_hovernet)."""
import numpy as np
import torch

from tests.cnn_common import hovernet_test_weights, rel_l2
from tiatoolbox_b200.models.architecture.hovernet import hovernet_forward_reference
from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan
from tiatoolbox_b200.synthetic import he_tiles

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sd = hovernet_test_weights("fast", 6)
import sys
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
plan = HoVerNetPlan(sd, batch=2, device="cuda:0", mode="fast", num_types=6, precision=prec)
print("precision", prec)
tiles = he_tiles(2, 256, 256, seed=0)
x = torch.from_numpy(tiles).cuda()
plan.forward(x)
torch.cuda.synchronize()
cap = {}
with torch.inference_mode():
    ref = hovernet_forward_reference({k: v.cuda() for k, v in sd.items()}, x.permute(0, 3, 1, 2).float(), "fast", 6, cap)


def cmp(name, got, r, win=None):
    r = r.permute(0, 2, 3, 1).float().cpu().numpy()
    got = got.float().cpu().numpy()
    if win is not None:
        lo, n = win
        got = got[:, lo:lo + n, lo:lo + n]
    print(f"{name:14s} shape {got.shape} rel_l2 {rel_l2(got, r):.3e}  max|ref| {np.abs(r).max():.3f}  finite {np.isfinite(got).all()}")


def val(name, c):
    t = plan.bufs[name]
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
# the shared scratch holds the LAST branch (hv)
cmp("hv.u3.conva", val("cb3", 256), cap["hv.u3.conva"], (1, 62))
cmp("hv.u3out", val("u3out", 512), cap["hv.u3out"], (9, 46))
cmp("hv.u2in", val("u2in", 512), cap["hv.u2in"])
cmp("hv.u2out", val("u2out", 256), cap["hv.u2out"], (5, 82))
cmp("hv.u1in", val("u1in_padded", 256), cap["hv.u1in"], (1, 164))
for b in ("tp", "np", "hv"):
    cmp(f"logits {b}", plan.logits[b], ref[b])
