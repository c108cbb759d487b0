"""Generate tests/golden/wsi_merge_golden.npz from the UNMODIFIED reference functions (oracle/ref_shim.load_wsi_merge):
  * `_get_tile_info` tile sets / flags for a list of slide shapes and margins,
  * `PatchExtractor.get_coordinates` for the HoVer-Net patch grid,
  * the surviving instance boxes of `_process_tile_mode`'s merge loop on seeded random tile contents (tests/wsi_common.scenario),
  * `_update_tile_based_predictions_array` on a seeded sequence of overlapping instance tiles.
Run in the build container (needs /root/reference):  python oracle/make_golden_wsi.py"""
from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from tests import wsi_common as wc  # noqa: E402

TILE_CASES = [((2500, 2150), 128), ((984, 984), 128), ((985, 500), 128), ((500, 985), 64), ((1968, 1968), 0), ((1969, 1000), None),
              ((3000, 900), 128), ((900, 3000), 128), ((5781, 5822), 128), ((300, 200), 128)]
MERGE_CASES = [(0, (2500, 2150), 128), (1, (3100, 1200), 128), (2, (1100, 3000), 64), (3, (1968, 1968), 0), (4, (4000, 3100), 128)]


def ref_tile_info(m, image_shape, tile_shape, pos, margin):
    self = SimpleNamespace(_ioconfig=SimpleNamespace(margin=margin, tile_shape=tile_shape, patch_output_shape=pos),
                           mask_padding=(0, 0, 0, 0), dataloader=SimpleNamespace(dataset=SimpleNamespace(mask_reader=None)))
    return m._get_tile_info(self, image_shape=image_shape, wsi_proc_shape=image_shape)


def paste_case(seed: int):
    """A 300x260 WSI instance map receiving 7 overlapping tiles with local ids 1..k."""
    rng = np.random.default_rng(seed)
    tiles = []
    for _ in range(7):
        x0, y0 = int(rng.integers(0, 220)), int(rng.integers(0, 200))
        w, h = int(rng.integers(40, 120)), int(rng.integers(40, 120))
        pred = np.zeros((h, w), np.int32)
        for k in range(1, int(rng.integers(2, 7))):
            cx, cy, r = int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(4, 14))
            yy, xx = np.ogrid[:h, :w]
            pred[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = k
        tiles.append((np.array([x0, y0, x0 + w, y0 + h]), pred))
    return tiles


def main():
    m = ref_shim.load_wsi_merge()
    data = {}
    for i, (shape, margin) in enumerate(TILE_CASES):
        sets = ref_tile_info(m, shape, (1024, 1024), (164, 164), margin)
        data[f"tile{i}.nsets"] = np.array(len(sets))
        for s, (boxes, flags) in enumerate(sets):
            data[f"tile{i}.boxes{s}"] = np.asarray(boxes).reshape(-1, 4)
            data[f"tile{i}.flags{s}"] = np.asarray(flags).reshape(-1, 4)
    ins, outs = m.PatchExtractor.get_coordinates(image_shape=(1000, 777), patch_input_shape=(256, 256), patch_output_shape=(164, 164),
                                                 stride_shape=(164, 164))
    data["coords.in"], data["coords.out"] = ins, outs
    for seed, shape, margin in MERGE_CASES:
        tiles = wc.scenario(seed, image_shape=shape, margin=margin, density=30)
        kept = wc.replay_reference(m, tiles, margin)
        data[f"merge{seed}.boxes"] = np.array([r["box"] for r in kept], dtype=np.int64).reshape(-1, 4)
        data[f"merge{seed}.prob"] = np.array([r["prob"] for r in kept], dtype=np.float64)
    for seed in (0, 1):
        wsi = np.zeros((260, 300), np.int32)
        info = ({"task_type": "nuclei_segmentation", "predictions": wsi, "info_dict": {}},)
        mx, trace = None, []
        for bounds, pred in paste_case(seed):
            post = ({"task_type": "nuclei_segmentation", "predictions": pred, "seg_type": "instance"},)
            info, mx = m._update_tile_based_predictions_array(post_process_output=post, wsi_info_dict=info, bounds=bounds, offset=(0, 0),
                                                              max_inst_value=mx)
            trace.append(int(mx))
        data[f"paste{seed}.wsi"] = info[0]["predictions"]
        data[f"paste{seed}.max"] = np.array(trace)
    out = ROOT / "tests" / "golden" / "wsi_merge_golden.npz"
    np.savez_compressed(out, **data)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
