"""Time the HOST side of the WSI merge replay (`WsiColumnStore.merge_tile` over all tiles + `to_info_dict`) on a synthetic tile sequence
with the statistics of BASELINE config 5 (23 288 x 23 124 px, 2 209 tiles, ~34 nuclei per 164 x 164 patch).  No GPU: the instance-map
paste is not part of this.  Usage: python tools/time_wsi_replay_host.py [--check]  (--check also runs the per-record store: slow)."""
from __future__ import annotations

import argparse
import time

import numpy as np

from tiatoolbox_b200.models.architecture import PRETRAINED_INFO
from tiatoolbox_b200.models.engine import wsi_merge as wm
from tiatoolbox_b200.models.engine.wsi import WsiTilePlan


def synth_tile(rng, bounds, density: float, verts: int = 18) -> dict:
    x0, y0, x1, y1 = (int(v) for v in bounds)
    w, h = x1 - x0, y1 - y0
    n = int(w * h * density)
    cx, cy = rng.uniform(0, w, n), rng.uniform(0, h, n)
    r = rng.uniform(4, 12, n)
    box = np.stack([np.clip(cx - r, 0, w), np.clip(cy - r, 0, h), np.clip(cx + r + 1, 0, w), np.clip(cy + r + 1, 0, h)], 1).astype(np.int64)
    ang = np.linspace(0, 2 * np.pi, verts, endpoint=False)
    xy = np.stack([cx[:, None] + r[:, None] * np.cos(ang), cy[:, None] + r[:, None] * np.sin(ang)], -1).astype(np.int32).reshape(-1, 2)
    return {"box": box, "centroid": np.stack([cx, cy], 1), "prob": rng.uniform(0.5, 1, n), "type": rng.integers(0, 6, n).astype(np.int64),
            "cont_xy": xy, "cont_off": np.arange(n + 1, dtype=np.int64) * verts}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--scale", type=float, default=1.0, help="linear scale of the slide (1.0 = config 5)")
    a = ap.parse_args()
    io = PRETRAINED_INFO["hovernet_fast-pannuke"]["ioconfig"]
    plan = WsiTilePlan((int(23124 * a.scale), int(23288 * a.scale)), io)
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    tiles = [synth_tile(rng, b, 34 / (164 * 164)) for b, _, _ in plan.tasks]
    print(f"{len(tiles)} tiles, {sum(len(t['box']) for t in tiles)} tile-level records (generated in {time.perf_counter() - t0:.1f} s)")
    store = wm.WsiColumnStore()
    t0 = time.perf_counter()
    for cols, (bounds, flag, mode) in zip(tiles, plan.tasks):
        store.merge_tile(cols, tile_bounds=bounds, tile_flag=flag, tile_mode=mode, margin=plan.margin, consume=not a.check)
    t1 = time.perf_counter()
    out = store.to_info_dict()
    t2 = time.perf_counter()
    print(f"merge_tile x {len(tiles)}: {t1 - t0:.2f} s; to_info_dict ({len(out['box'])} instances): {t2 - t1:.2f} s")
    if a.check:
        ref = wm.WsiInstanceStore()
        for cols, (bounds, flag, mode) in zip(tiles, plan.tasks):
            n = len(cols["box"])
            d = {i + 1: {"box": cols["box"][i], "centroid": cols["centroid"][i], "prob": cols["prob"][i], "type": int(cols["type"][i]),
                         "contours": cols["cont_xy"][cols["cont_off"][i]:cols["cont_off"][i + 1]]} for i in range(n)}
            ref.merge_tile(d, tile_bounds=bounds, tile_flag=flag, tile_mode=mode, margin=plan.margin)
        want = ref.to_info_dict()
        assert len(want["box"]) == len(out["box"])
        for key in want:
            assert all(np.array_equal(x, y) for x, y in zip(want[key], out[key])), key
        print("equal to the per-record store")


if __name__ == "__main__":
    main()
