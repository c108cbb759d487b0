import torch, sys
import torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from tiatoolbox_b200.models import ops
from tiatoolbox_b200.models.plan import pack_weight
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
def run(k, cin, cout, H=24, W=26, nimg=2, split=True, seed=0):
    torch.manual_seed(seed)
    xf = torch.randn(nimg, H, W, cin, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    hi = xf.half(); lo = (xf - hi.float()).half()
    x = torch.cat([hi, lo], dim=-1).contiguous() if split else hi.contiguous()
    wp = pack_weight(w, split=split, device=dev)
    M = nimg * H * W
    taps = [(kh - k // 2) * W + (kw - k // 2) for kh in range(k) for kw in range(k)]
    out32 = torch.zeros(nimg, H, W, cout, device=dev)
    h = k // 2
    op = ops.ConvOp(x=x, x_rows=M, x_ld=x.shape[-1], cin=cin, x_lo_off=cin if split else 0, w=wp, taps=taps, M=M, N=cout, grid=(H, W),
                    window=(h, H - h, h, W - h), dst_geom=(H, W, 0, 0), out32=out32, ld32=cout)
    op(); torch.cuda.synchronize()
    xin = (hi.float() + lo.float()) if split else hi.float()
    ref = F.conv2d(xin.double().permute(0, 3, 1, 2), w.double()).permute(0, 2, 3, 1).float()
    got = out32[:, h:H - h, h:W - h]
    e = ((got - ref).norm() / ref.norm()).item()
    print(f"k={k} cin={cin} cout={cout} split={split}: rel_l2 {e:.3e}")
for k, cin, cout in [(3, 64, 64), (5, 64, 64), (3, 1024, 256), (5, 1024, 256), (5, 256, 256), (7, 128, 128), (5, 512, 128), (5, 128, 32)]:
    run(k, cin, cout)
run(5, 1024, 256, split=False)
