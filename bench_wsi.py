"""BASELINE config 5 harness: Vahadane stain-norm -> HoVer-Net (hovernet_fast) end to end on a synthetic WSI, N GPUs, gather of the instance map.

    python bench_wsi.py --patches 2000                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench_wsi.py --patches 20000

`bench.py` stays the driver's contract (the metric's headline configuration: Macenko + HoVer-Net patches).  This script times the WSI-mode
public call — `NucleusInstanceSegmentor.run(slide, patch_mode=False, return_predictions=True)` with the Vahadane normaliser installed as
`preproc_func` — on a virtual slide generated tile-wise from a seed: host slide in, merged instance table + int32
instance map out, i.e. everything (upload, patches, stain, network, stitching, tile post-processing, contours, gather, merge replay) is inside
the timed region.  Wall clock around a barrier + synchronize on both sides, max over ranks; prints one JSON line on rank 0.
Status: the data flow is parity-green on the B200 at small size (tests/test_wsi_gpu.py); this harness itself has not been run yet (round 1's GPU
budget was spent before it was written)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def synthetic_slide(n_patches: int, seed: int = 0) -> np.ndarray:
    """H&E-like slide whose 164-px output grid has about `n_patches` cells (141 x 142 for 20 000), built from seeded 256-px tiles."""
    from tiatoolbox_b200.synthetic import he_tiles

    nx = max(1, int(math.floor(math.sqrt(n_patches))))
    ny = max(1, int(math.ceil(n_patches / nx)))
    h, w = ny * 164, nx * 164
    base = he_tiles(16, 256, 256, seed=seed)
    rng = np.random.default_rng(seed)
    slide = np.empty((h, w, 3), np.uint8)
    for y in range(0, h, 256):
        for x in range(0, w, 256):
            t = base[int(rng.integers(0, len(base)))]
            slide[y:y + 256, x:x + 256] = t[: min(256, h - y), : min(256, w - x)]
    return slide


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--method", default="vahadane", choices=["vahadane", "macenko", "reinhard", "ruifrok"])
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    g.build()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bench import _hovernet_weights
    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_target
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    slide = synthetic_slide(args.patches)
    norm = get_normalizer(args.method)
    norm.device = local
    norm.fit(he_target(256, 256))
    eng = NucleusInstanceSegmentor("hovernet_fast-pannuke", batch_size=args.batch, device=f"cuda:{local}")
    eng.model.load_state_dict(_hovernet_weights())
    eng.model.preproc_func = norm.transform

    def once():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.run(slide, patch_mode=False, return_predictions=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter() - t0, out

    once()                                  # warm-up: plan build, CUDA graph capture, pinned buffers
    times, out = [], None
    for _ in range(args.repeats):
        dt, out = once()
        times.append(dt)
    t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = len(out["coordinates"])
        print(json.dumps({
            "metric": "patches/sec (256x256x3) stain-norm + HoVer-Net end to end on a synthetic WSI (config 5)", "value": n / float(t.item()),
            "unit": "patches/s", "n_gpus": world, "seconds_per_slide": float(t.item()), "higher_is_better": True, "scaling": "strong",
            "data": "synthetic", "dtype": "f16x3",
            "config": {"workload": f"{args.method}_hovernet_wsi", "slide": list(slide.shape), "patches": n, "batch": args.batch,
                       "instances": int(len(out.get("box", []))), "timed": "host slide -> merged instance table + int32 instance map (wall clock)"},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
