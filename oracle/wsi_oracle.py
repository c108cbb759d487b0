"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's WSI tile-mode post-processing bookkeeping.

Follows tiatoolbox/models/engine/multi_task_segmentor.py:
  _get_tile_info                        :1362-1554   tile grid + vertical / horizontal strips + cross-sections and their removal flags
  _build_tile_tasks                     :3215-3266   processing order
  _get_sel_indices_margin_lines         :2944-3020   which of a tile's instances are dropped
  _get_margin_lines                     :3023-3038
  _process_instance_predictions         :2833-2933   + _move_tile_space_to_wsi_space :3041-3058, _compute_info_dict_for_merge :3269
  _update_tile_based_predictions_array  :3160-3212
  _process_tile_mode                    :1230-1279   (the merge loop; dask / zarr / tqdm dropped)
and tiatoolbox/tools/patchextraction.py:488-613 (get_coordinates).

The reference tests its geometry with shapely (`shapely.box`, `STRtree.query`, `.contains`); shapely is a pinned dependency that is
absent from this image.  Every geometry involved is an axis-aligned rectangle, so the restatement is scalar interval logic, written
out per instance (deliberately unvectorised: the product's `wsi_merge.py` is the vectorised one).

Parity status: PINNED against the unmodified reference functions executed through `oracle/ref_shim.load_wsi_merge()` with shapely's
two calls restated (tests/test_oracle_wsi.py, and the committed `tests/golden/wsi_merge_golden.npz` made by `oracle/make_golden_wsi.py`);
"parity unpinned" only with respect to shapely itself (rectangle `intersects` / `contains` semantics are restated, not executed).
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

import math

import numpy as np


# ---------------------------------------------------------------- rectangles
def _touch(a, b) -> bool:
    """closed rectangles share at least one point (what an STRtree envelope query reports)"""
    return not (a[2] < b[0] or b[2] < a[0] or a[3] < b[1] or b[3] < a[1])


def _within(inner, outer) -> bool:
    """outer.contains(inner) for rectangles; an outer without area contains nothing"""
    if not (outer[2] > outer[0] and outer[3] > outer[1]):
        return False
    return outer[0] <= inner[0] and outer[1] <= inner[1] and inner[2] <= outer[2] and inner[3] <= outer[3]


# ---------------------------------------------------------------- patch / tile grids
def get_coordinates(image_shape, patch_input_shape, patch_output_shape, stride_shape):
    """patchextraction.py:488-613 with both within_bound flags False: (input boxes, output boxes), x fastest."""
    iw, ih = int(image_shape[0]), int(image_shape[1])
    inputs, outputs = [], []
    dx = (int(patch_input_shape[0]) - int(patch_output_shape[0])) // 2
    dy = (int(patch_input_shape[1]) - int(patch_output_shape[1])) // 2
    for gy in range(math.ceil(ih / int(stride_shape[1]))):
        for gx in range(math.ceil(iw / int(stride_shape[0]))):
            ox, oy = gx * int(stride_shape[0]), gy * int(stride_shape[1])
            outputs.append([ox, oy, ox + int(patch_output_shape[0]), oy + int(patch_output_shape[1])])
            inputs.append([ox - dx, oy - dy, ox - dx + int(patch_input_shape[0]), oy - dy + int(patch_input_shape[1])])
    return np.array(inputs, dtype=np.int64).reshape(-1, 4), np.array(outputs, dtype=np.int64).reshape(-1, 4)


def get_tile_info(image_shape, tile_shape, patch_output_shape, margin):
    """:1362-1554 (no mask reader).  Returns [[boxes, flags], ...] — one set if the image fits one tile, else four."""
    margin = 0 if margin is None else int(margin)
    tw = int(math.floor(tile_shape[0] / patch_output_shape[0]) * patch_output_shape[0])
    th = int(math.floor(tile_shape[1] / patch_output_shape[1]) * patch_output_shape[1])
    w, h = int(image_shape[0]), int(image_shape[1])
    grid = get_coordinates((w, h), (tw, th), (tw, th), (tw, th))[1]
    if w <= tw and h <= th:
        return [[grid, np.zeros((len(grid), 4), dtype=np.int32)]]
    borders = [(0, 0, w, 0), (0, h, w, h), (0, 0, 0, h), (w, 0, w, h)]

    def clear_at_borders(boxes, flags):
        for i, bx in enumerate(boxes):
            for side, border in enumerate(borders):
                if _touch(bx, border):
                    flags[i][side] = 0
        return flags

    gflags = clear_at_borders(grid, [[1, 1, 1, 1] for _ in grid])
    sets = [[grid, np.array(gflags, dtype=np.int32).reshape(-1, 4)]]
    vert = [[x1 - margin, y0, x1 + margin, y1] for (x0, y0, x1, y1), f in zip(grid.tolist(), gflags) if f[3]]
    vflags = clear_at_borders(vert, [[1, 1, 0, 0] for _ in vert])
    sets.append([np.array(vert, dtype=np.int64).reshape(-1, 4), np.array(vflags, dtype=np.int32).reshape(-1, 4)])
    horz = [[x0, y1 - margin, x1, y1 + margin] for (x0, y0, x1, y1), f in zip(grid.tolist(), gflags) if f[1]]
    hflags = clear_at_borders(horz, [[0, 0, 1, 1] for _ in horz])
    sets.append([np.array(horz, dtype=np.int64).reshape(-1, 4), np.array(hflags, dtype=np.int32).reshape(-1, 4)])
    cross = [[x1 - 2 * margin, y1 - 2 * margin, x1 + 2 * margin, y1 + 2 * margin]
             for (x0, y0, x1, y1), f in zip(grid.tolist(), gflags) if f[1] and f[3]]
    sets.append([np.array(cross, dtype=np.int64).reshape(-1, 4), np.ones((len(cross), 4), dtype=np.int32)])
    return sets


def build_tile_tasks(sets):
    return [(b, flags[i], mode) for mode, (boxes, flags) in enumerate(sets) for i, b in enumerate(boxes)]


# ---------------------------------------------------------------- instance merge
def tile_drop_list(boxes, tile_shape, tile_flag, tile_mode, margin):
    """:2944-3020 — indices (into the tile's instance list) that are discarded."""
    margin = 0 if margin is None else int(margin)
    w, h = int(tile_shape[0]), int(tile_shape[1])
    edge = [(0, 0, w, 1), (0, h - 1, w, h), (0, 0, 1, h), (w - 1, 0, w, h)]
    band = [(0, 0, w, margin), (0, h - margin, w, h), (0, 0, margin, h), (w - margin, 0, w, h)]
    dropped = []
    for i, bx in enumerate(boxes):
        if tile_mode in (0, 3):
            chosen = [band[s] for s in range(4) if tile_flag[s] or tile_mode == 3]
            if any(_within(bx, region) for region in chosen):
                dropped.append(i)
        elif tile_mode in (1, 2):
            chosen = [band[s] if tile_flag[s] else edge[s] for s in range(4)]
            if any(_touch(bx, region) for region in chosen):
                dropped.append(i)
        else:
            raise ValueError(f"Unknown tile mode {tile_mode}.")
    return dropped


def replay_merge(tiles, margin):
    """The merge loop of `_process_tile_mode` for one instance task.  `tiles` = [(bounds, flag, mode, [record, ...])] in processing
    order, records in tile space with at least a "box".  Returns the surviving records in WSI space, in dictionary order."""
    margin_ = 0 if margin is None else int(margin)
    acc: dict[int, dict] = {}
    serial = 0
    for bounds, flag, mode, records in tiles:
        if len(records) == 0:
            continue
        tlx, tly = int(bounds[0]), int(bounds[1])
        w, h = int(bounds[2]) - tlx, int(bounds[3]) - tly
        dropped = set(tile_drop_list([r["box"] for r in records], (w, h), flag, mode, margin))
        evict = []
        if mode == 3:
            m = margin_
            lines = [(m + tlx, m + tly, w - m + tlx, m + tly), (m + tlx, h - m + tly, w - m + tlx, h - m + tly),
                     (m + tlx, m + tly, m + tlx, h - m + tly), (w - m + tlx, m + tly, w - m + tlx, h - m + tly)]
            for key, rec in acc.items():
                if any(_touch(rec["box"], ln) for ln in lines):
                    evict.append(key)
        for i, rec in enumerate(records):
            if i in dropped:
                continue
            moved = dict(rec)
            for key, shift in (("box", [tlx, tly, tlx, tly]), ("centroid", [tlx, tly]), ("contours", [tlx, tly])):
                if key in rec:   # the reference adds in place, so the dtype of the record's array survives
                    moved[key] = np.array(rec[key], copy=True)
                    moved[key] += np.array(shift)
            acc[serial] = moved
            serial += 1
        for key in evict:
            acc.pop(key, None)
    return list(acc.values())


def paste_instance_tile(wsi_pred, tile_pred, bounds, max_inst_value):
    """:3160-3212 for an instance map, in place; returns the new running `max_inst_value`."""
    x0, y0, x1, y1 = (int(v) for v in bounds)
    H, W = wsi_pred.shape
    x1, y1 = min(x1, W), min(y1, H)
    new = tile_pred[: y1 - y0, : x1 - x0]
    prev = wsi_pred[y0:y1, x0:x1]
    base = 0 if max_inst_value is None else max_inst_value
    clash = set(np.unique(new[(new > 0) & (prev > 0)]).tolist()) - {0}
    merged = prev.copy()
    kept_any = False
    for yy in range(new.shape[0]):
        for xx in range(new.shape[1]):
            v = int(new[yy, xx])
            if v > 0 and v not in clash:
                merged[yy, xx] = v + base
                kept_any = True
    wsi_pred[y0:y1, x0:x1] = merged
    return (int(merged.max()) + base) if kept_any else base
