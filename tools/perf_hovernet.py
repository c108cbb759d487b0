"""Manual perf probe (GPU box): python -m tools.perf_hovernet [batch] [precision]"""
import sys
import time

import torch

from tests.cnn_common import hovernet_test_weights
from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan
from tiatoolbox_b200.synthetic import he_tiles

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
t0 = time.time()
sd = hovernet_test_weights("fast", 6)
t1 = time.time()
plan = HoVerNetPlan(sd, batch=batch, device="cuda:0", mode="fast", num_types=6, precision=prec)
torch.cuda.synchronize()
t2 = time.time()
x = torch.from_numpy(he_tiles(4, 256, 256, seed=0)).cuda().repeat((batch + 3) // 4, 1, 1, 1)[:batch].contiguous()
plan.forward(x)
torch.cuda.synchronize()
t3 = time.time()
plan.capture()
t4 = time.time()
print(f"weights {t1-t0:.1f}s plan {t2-t1:.1f}s first fwd {t3-t2:.2f}s capture {t4-t3:.2f}s")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    plan.forward(x)
s.record()
n = 10
for _ in range(n):
    plan.forward(x)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
print(f"hovernet_fast B={batch} {prec}: {ms:.3f} ms/step  {batch / ms * 1e3:.1f} patches/s  useful {plan.flops / ms / 1e9:.1f} TFLOP/s  "
      f"mma {plan.mma_flops / ms / 1e9:.1f} TFLOP/s  launches {plan.num_launches}  buffers {plan.bytes_allocated() / 1e9:.2f} GB")
if len(sys.argv) > 3:  # per-op timing
    torch.cuda.synchronize()
    times = []
    for i, op in enumerate(plan.ops):
        s.record()
        for _ in range(3):
            op()
        e.record()
        torch.cuda.synchronize()
        fl = getattr(op, "flops", 0.0)
        mm = getattr(op, "mma_flops", 0.0)
        d = getattr(op, "desc", None)
        info = f"M={d.M} N={d.N} taps={d.taps} cin={d.cin}" if d is not None else "aux"
        times.append((s.elapsed_time(e) / 3, i, info, fl, mm))
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump([{"ms": t, "op": i, "info": info, "flops": fl, "mma": mm, "name": getattr(plan.ops[i], "__name__", type(plan.ops[i]).__name__)}
               for t, i, info, fl, mm in times], open(f"gpurun_out/ops_B{batch}_{prec}.json", "w"))
    tot = sum(t[0] for t in times)
    print(f"sum of per-op times {tot:.3f} ms")
    for t, i, info, fl, mm in sorted(times, reverse=True)[:40]:
        print(f"  op{i:3d} {t*1e3:8.1f} us  {100*t/tot:5.1f}%  {info:36s} useful {fl/t/1e9 if t else 0:7.1f} TF/s  mma {mm/t/1e9 if t else 0:7.1f} TF/s")
