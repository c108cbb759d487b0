"""CUDA-event timings of the HoVer-Net post-processing (tiab_hovernet_postproc) per flood implementation on SURVEY §8d config-4 head maps
(~40 nuclei per 164-px map).  Prints one JSON line per (implementation, batch).   python profiles/time_postproc.py [impl ...]"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main() -> None:
    import __graft_entry__ as g

    g.build()
    from tiatoolbox_b200.models import postproc as pp
    from tiatoolbox_b200.synthetic import hovernet_head_maps

    impls = [int(a) for a in sys.argv[1:]] or [1, 3]
    for n in (16, 64):
        maps = [torch.from_numpy(a).cuda() for a in hovernet_head_maps(n, 164, seed=0)]
        ref = None
        for impl in impls:
            pp.set_watershed_impl(impl)
            for _ in range(3):
                inst, table = pp.proc_np_hv_batch(*maps)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                inst, table = pp.proc_np_hv_batch(*maps)
            b.record()
            torch.cuda.synchronize()
            same = None if ref is None else bool(torch.equal(inst, ref))
            if ref is None:
                ref = inst.clone()
            print(json.dumps({"flood_impl": impl, "maps": n, "postproc_ms": a.elapsed_time(b) / 10, "nuclei": int(table["count"].sum()),
                              "equal_to_first_impl": same}), flush=True)
    pp.set_watershed_impl(4)


if __name__ == "__main__":
    main()
