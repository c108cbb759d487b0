"""Error study (CPU emulation): which layers of HoVer-Net need the split-fp16 ("fp16x3") product for the 1e-3 end-to-end bar?
The encoder is a 57-conv-deep pre-activation ResNet (operand rounding error accumulates with depth); every decoder branch is only
~8 convolutions from the output.  Emulated per layer: operands rounded to fp16 (hi) and fp16 residuals (lo), products / accumulation in fp64.
  policy A  every layer fp16x3 (the r1 default)
  policy B  decoder `conva` convolutions (3x3 1024->256, 512->128, 256->64: 36 % of all FLOPs) single fp16 pass, rest fp16x3
  policy C  the whole decoder single fp16 pass, encoder fp16x3
  policy D  decoder conva: A_hi*(W_hi + W_lo) (activation hi plane only, weights split), rest fp16x3
  policy E  every layer single fp16
python profiles/studies/decoder_precision_study.py [seed]"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.nn.functional as F
from tests.cnn_common import hovernet_test_weights, rel_l2
import oracle.torch_reference as hn
from tiatoolbox_b200.synthetic import he_tiles

real_conv = F.conv2d
N_ENCODER_CONVS = 54          # conv0 + d0 (10) + d1 (13) + d2 (19) + d3 (10) + conv_bot
def f16(x): return x.half().float()
STATE = {"i": 0, "policy": "fp32"}
def conv_emul(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
    i = STATE["i"]; STATE["i"] += 1
    pol = STATE["policy"]
    if pol == "fp32":
        return real_conv(x, w, b, stride, padding, dilation, groups)
    decoder = i >= N_ENCODER_CONVS
    conva = decoder and w.shape[2] == 3 and w.shape[1] in (1024, 512, 256) and groups == 1
    mode = "x3"
    if pol == "B" and conva: mode = "x1"
    if pol == "C" and decoder: mode = "x1"
    if pol == "D" and conva: mode = "hiA"
    if pol == "E": mode = "x1"
    xh, wh = f16(x), f16(w)
    xl, wl = f16(x - xh), f16(w - wh)
    c = lambda a, bb: real_conv(a.double(), bb.double(), None, stride, padding, dilation, groups).float()
    if mode == "x1": y = c(xh, wh)
    elif mode == "hiA": y = c(xh, wh) + c(xh, wl)
    else: y = c(xh, wh) + c(xl, wh) + c(xh, wl)
    if b is not None: y = y + b.view(1, -1, 1, 1)
    return y
hn.F.conv2d = conv_emul
sd = hovernet_test_weights("fast", 6)
x = torch.from_numpy(he_tiles(1, 256, 256, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 0)).permute(0, 3, 1, 2).float()
res = {}
for pol in ["fp32", "A", "B", "D", "C", "E"]:
    STATE["policy"], STATE["i"] = pol, 0
    t = time.time()
    with torch.inference_mode():
        out = hn.hovernet_forward_reference(sd, x, "fast", 6)
    res[pol] = {k: v.numpy() for k, v in out.items()}
    if pol != "fp32":
        r = res["fp32"]
        print(f"policy {pol}", {k: f"rel_l2 {rel_l2(res[pol][k], r[k]):.2e} max/max {np.abs(res[pol][k]-r[k]).max()/np.abs(r[k]).max():.2e}" for k in out},
              f"{time.time()-t:.0f}s convs={STATE['i']}", flush=True)
