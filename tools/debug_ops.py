"""Manual debug probe (GPU box): run a HoVer-Net plan op by op with a sync after each, report the first failing launch."""
import sys

import torch

from tests.cnn_common import hovernet_test_weights
from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan
from tiatoolbox_b200.synthetic import he_tiles

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
sd = hovernet_test_weights("fast", 6)
plan = HoVerNetPlan(sd, batch=batch, device="cuda:0", mode="fast", num_types=6, precision=prec)
x = torch.from_numpy(he_tiles(4, 256, 256, seed=0)).cuda().repeat((batch + 3) // 4, 1, 1, 1)[:batch].contiguous()
plan.inp.copy_(x)
torch.cuda.synchronize()
for i, op in enumerate(plan.ops):
    d = getattr(op, "desc", None)
    info = (f"M={d.M} N={d.N} taps={d.taps} cin={d.cin} pre={bool(d.pre_scale)} f32={d.x_is_f32} split={d.split} lo={d.x_lo_off} "
            f"win=({d.y0},{d.y1},{d.x0},{d.x1}) grid=({d.Hin},{d.Win}) out16={bool(d.out16)} out32={bool(d.out32)} out16b={bool(d.out16b)} "
            f"res32={bool(d.res32)}") if d is not None else "aux"
    try:
        op()
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        print(f"FAILED at op {i}: {info}\n  {exc}")
        break
else:
    print("all ops ran")
