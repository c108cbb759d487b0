"""WSI tile-mode bookkeeping of MultiTaskSegmentor (mirror of tiatoolbox/models/engine/multi_task_segmentor.py:1362-1554, 2833-3296).

The reference post-processes a WSI-scale probability map tile by tile — a non-overlapping tile grid, then vertical strips, horizontal
strips and cross-sections straddling the tile seams — and merges the per-tile instance dictionaries *in that order*: instances that sit
wholly inside (grid / cross tiles) or that touch (strips) a flagged margin of their tile are dropped, and a cross tile evicts the
already accumulated instances its inner margin lines cut.  On the B200 path the per-tile watershed runs on the device
(`postproc.proc_np_hv_batch`) and on any rank; what is left is this order-dependent index arithmetic, replayed on rank 0 after the
gather.  The reference drives it through shapely boxes and a fresh STRtree per tile; all geometries are axis-aligned
rectangles, so here it is closed-interval arithmetic on int64 arrays (no shapely on this path).

Every function cites the reference function it replaces; `tests/test_oracle_wsi.py` replays random tile sequences through the unmodified reference functions, the oracle and this module.
"""
from __future__ import annotations

import numpy as np

from tiatoolbox_b200.tools.patchextraction import PatchExtractor

INFO_KEYS = ("box", "centroid", "contours", "prob", "type")


def boxes_intersect(boxes: np.ndarray, query) -> np.ndarray:
    """STRtree.query(box) without predicate: envelopes intersect, closed intervals (touching counts)."""
    b = np.asarray(boxes).reshape(-1, 4)
    x0, y0, x1, y1 = (min(query[0], query[2]), min(query[1], query[3]), max(query[0], query[2]), max(query[1], query[3]))
    return (b[:, 0] <= x1) & (b[:, 2] >= x0) & (b[:, 1] <= y1) & (b[:, 3] >= y0)


def box_contains(outer, boxes: np.ndarray) -> np.ndarray:
    """shapely `outer.contains(box)` for rectangles: closed inclusion in an `outer` that has area (a line contains no rectangle)."""
    b = np.asarray(boxes).reshape(-1, 4)
    if outer[2] <= outer[0] or outer[3] <= outer[1]:
        return np.zeros(len(b), dtype=bool)
    return (outer[0] <= b[:, 0]) & (outer[1] <= b[:, 1]) & (b[:, 2] <= outer[2]) & (b[:, 3] <= outer[3])


def _shifted(v, shift) -> np.ndarray:
    v = np.asarray(v)
    return (v + shift).astype(v.dtype)


def get_tile_info(image_shape, tile_shape, patch_output_shape, margin) -> list[list[np.ndarray]]:
    """`MultiTaskSegmentor._get_tile_info` (:1362-1554) without a mask reader: tile placement + removal flags [top, bottom, left, right]
    for the four tile sets (grid, vertical strips, horizontal strips, cross-sections).  `image_shape` is (width, height)."""
    margin = 0 if margin is None else int(margin)
    pos = np.asarray(patch_output_shape)
    tile_shape = (np.floor(np.asarray(tile_shape) / pos) * pos).astype(np.int32)
    image_shape = np.asarray(image_shape)
    boxes = PatchExtractor.get_coordinates(image_shape=image_shape, patch_input_shape=tile_shape, patch_output_shape=tile_shape,
                                           stride_shape=tile_shape)[1]
    if np.all(image_shape <= tile_shape):
        return [[boxes, np.zeros([boxes.shape[0], 4], dtype=np.int32)]]
    w, h = (int(v) for v in image_shape)
    edges = ((0, 0, w, 0), (0, h, w, h), (0, 0, 0, h), (w, 0, w, h))   # top, bottom, left, right image borders

    def unset_removal_flag(bx: np.ndarray, flag: np.ndarray) -> np.ndarray:
        for idx, edge in enumerate(edges):
            flag[boxes_intersect(bx, edge), idx] = 0
        return flag

    flag = unset_removal_flag(boxes, np.ones([boxes.shape[0], 4], dtype=np.int32))
    info = [[boxes, flag]]
    x0, y0, x1, y1 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    # vertical strips over the right edge of every tile whose right side is removed; their top / bottom are removed
    s = np.nonzero(flag[:, 3])[0]
    vb = np.stack([x1[s] - margin, y0[s], x1[s] + margin, y1[s]], axis=-1)
    vf = np.zeros([len(s), 4], dtype=np.int32)
    vf[:, [0, 1]] = 1
    info.append([vb, unset_removal_flag(vb, vf)])
    # horizontal strips over the bottom edge; their left / right are removed
    s = np.nonzero(flag[:, 1])[0]
    hb = np.stack([x0[s], y1[s] - margin, x1[s], y1[s] + margin], axis=-1)
    hf = np.zeros([len(s), 4], dtype=np.int32)
    hf[:, [2, 3]] = 1
    info.append([hb, unset_removal_flag(hb, hf)])
    # cross-sections around the bottom-right corner; all sides removed
    s = np.nonzero(flag[:, 1] * flag[:, 3])[0]
    cb = np.stack([x1[s] - 2 * margin, y1[s] - 2 * margin, x1[s] + 2 * margin, y1[s] + 2 * margin], axis=-1)
    info.append([cb, np.ones([len(s), 4], dtype=np.int32)])
    return info


def build_tile_tasks(tile_info_sets) -> list[tuple[np.ndarray, np.ndarray, int]]:
    """`_build_tile_tasks` (:3215-3266): [(tile_bounds, tile_flag, tile_mode)] in processing order (set by set)."""
    return [(bounds, set_flags[i], set_idx) for set_idx, (set_bounds, set_flags) in enumerate(tile_info_sets)
            for i, bounds in enumerate(set_bounds)]


def select_tile_removals(inst_boxes: np.ndarray, tile_shape, tile_flag, tile_mode: int, margin) -> np.ndarray:
    """`_get_sel_indices_margin_lines` (:2944-3020), first result: bool mask of the tile's instances to drop (boxes in tile space)."""
    if tile_mode not in (0, 1, 2, 3):
        msg = f"Unknown tile mode {tile_mode}."
        raise ValueError(msg)
    margin = 0 if margin is None else int(margin)
    width, height = (int(v) for v in tile_shape)
    boundary_lines = ((0, 0, width, 1), (0, height - 1, width, height), (0, 0, 1, height), (width - 1, 0, width, height))
    margin_boxes = ((0, 0, width, margin), (0, height - margin, width, height), (0, 0, margin, height), (width - margin, 0, width, height))
    b = np.asarray(inst_boxes).reshape(-1, 4)
    drop = np.zeros(len(b), dtype=bool)
    if tile_mode in (0, 3):
        for idx, mb in enumerate(margin_boxes):
            if tile_flag[idx] or tile_mode == 3:  # noqa: PLR2004
                drop |= box_contains(mb, b)
        return drop
    for idx, flag in enumerate(tile_flag):
        drop |= boxes_intersect(b, margin_boxes[idx] if flag else boundary_lines[idx])
    return drop


def margin_lines(margin, tile_shape, tile_tl) -> np.ndarray:
    """`_get_margin_lines` (:3023-3038): the four inner margin lines of a tile in WSI space as (4, 4) [x0, y0, x1, y1] rows."""
    margin = 0 if margin is None else int(margin)
    width, height = (int(v) for v in tile_shape)
    lines = np.array([[margin, margin, width - margin, margin], [margin, height - margin, width - margin, height - margin],
                      [margin, margin, margin, height - margin], [width - margin, margin, width - margin, height - margin]], dtype=np.int64)
    return lines + np.concatenate([np.asarray(tile_tl), np.asarray(tile_tl)])[None]


class WsiInstanceStore:
    """The accumulated `wsi_info_dict[task]["info_dict"]` of `_process_tile_mode` (:1230-1279): an insertion-ordered set of instance
    records with evictions.  Boxes are mirrored in a growing int64 array so that a cross tile's eviction query is one vectorised
    interval test instead of an STRtree rebuilt over every accumulated instance (`_compute_info_dict_for_merge` :3269-3296)."""

    def __init__(self) -> None:
        self.records: list[dict] = []
        self._boxes = np.empty((1024, 4), dtype=np.int64)
        self._alive = np.zeros(1024, dtype=bool)
        self.n = 0

    def __len__(self) -> int:
        return int(self._alive[: self.n].sum())

    def _append(self, rec: dict) -> None:
        if self.n == len(self._boxes):
            self._boxes = np.concatenate([self._boxes, np.empty_like(self._boxes)])
            self._alive = np.concatenate([self._alive, np.zeros_like(self._alive)])
        self._boxes[self.n] = rec["box"]
        self._alive[self.n] = True
        self.records.append(rec)
        self.n += 1

    def merge_tile(self, inst_dict: dict, *, tile_bounds, tile_flag, tile_mode: int, margin) -> tuple[int, int]:
        """`_compute_info_dict_for_merge` + the update / pop of `_process_tile_mode` for one tile.  `inst_dict` = {k: record} in tile
        space (record keys: box [x0, y0, x1, y1], centroid, contours, prob, type), in the order of the tile's info_dict.
        Returns (#added, #evicted)."""
        if len(inst_dict) == 0:   # :2903-2904 — an empty cross tile evicts nothing either
            return 0, 0
        tile_bounds = np.asarray(tile_bounds)
        tile_tl = tile_bounds[:2]
        tile_shape = tile_bounds[2:] - tile_bounds[:2]
        recs = list(inst_dict.values())
        boxes = np.array([r["box"] for r in recs]).reshape(-1, 4)
        drop = select_tile_removals(boxes, tile_shape, tile_flag, tile_mode, margin)
        evicted = 0
        if tile_mode == 3:  # noqa: PLR2004
            hit = np.zeros(self.n, dtype=bool)
            for line in margin_lines(margin, tile_shape, tile_tl):
                hit |= boxes_intersect(self._boxes[: self.n], line)
            hit &= self._alive[: self.n]
            evicted = int(hit.sum())
            self._alive[: self.n][hit] = False
        shift4 = np.concatenate([tile_tl, tile_tl])
        added = 0
        for rec, d in zip(recs, drop):
            if d:
                continue
            moved = dict(rec)    # `_move_tile_space_to_wsi_space` (:3041-3058): in-place `+=`, i.e. every array keeps its dtype
            moved["box"] = _shifted(rec["box"], shift4)
            if "centroid" in rec:
                moved["centroid"] = _shifted(rec["centroid"], tile_tl)
            moved["contours"] = _shifted(rec["contours"], tile_tl)
            self._append(moved)
            added += 1
        return added, evicted

    def alive_records(self) -> list[dict]:
        return [r for r, a in zip(self.records, self._alive[: self.n]) if a]

    def to_info_dict(self, offset=(0, 0)) -> dict:
        """`_inst_dict_for_dask_processing` (:1289-1328): {key: object array over the surviving instances, insertion order}; the mask
        offset is added to box / centroid / contours (`apply_coordinate_offset` :3760-3826).  No instances -> {} like the reference."""
        recs = self.alive_records()
        if not recs:
            return {}
        dx, dy = (int(v) for v in offset)
        out = {}
        for key in recs[0]:
            col = np.empty(len(recs), dtype=object)
            for i, r in enumerate(recs):
                v = r[key]
                if key in ("centroid", "box", "contours") and not dx == dy == 0:
                    v = np.asarray(v)
                    shift = np.array([dx, dy, dx, dy]) if (v.size == 4 and v.ndim == 1) else np.array([dx, dy])  # noqa: PLR2004
                    v = (v + shift).astype(v.dtype)
                col[i] = v
            out[key] = col
        return out


def columns_from_records(recs) -> dict:
    """A tile's instances as COLUMNS (what the device table already is): box (n, 4) int64, centroid (n, 2) float64, prob (n,) float64,
    type (n,) int64 with -1 = None, contours as one (V, 2) int32 vertex array + (n + 1,) offsets.  `recs`: iterable of record dicts."""
    recs = list(recs)
    n = len(recs)
    cont = [np.asarray(r["contours"], dtype=np.int32).reshape(-1, 2) for r in recs]
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        off[1:] = np.cumsum([len(c) for c in cont])
    return {
        "box": np.array([r["box"] for r in recs], dtype=np.int64).reshape(n, 4),
        "centroid": np.array([r["centroid"] for r in recs], dtype=np.float64).reshape(n, 2),
        "prob": np.array([np.nan if r.get("prob") is None else r["prob"] for r in recs], dtype=np.float64),
        "type": np.array([-1 if r.get("type") is None else r["type"] for r in recs], dtype=np.int64),
        "cont_xy": np.concatenate(cont) if n else np.zeros((0, 2), np.int32),
        "cont_off": off,
    }


class WsiColumnStore:
    """`WsiInstanceStore` for columnar tiles: the same order-dependent merge (`_compute_info_dict_for_merge` :3269-3296 + the update / pop
    of `_process_tile_mode` :1230-1279), but a tile arrives as arrays (`columns_from_records` layout, tile space) and every step — the
    margin test, the shift to WSI space, the cross tile's eviction — is one vectorised operation per tile instead of Python work per
    nucleus.  At 20 000 patches a slide has ~2.5 M tile-level records (680 k survive): the per-record version spent longer on them than the
    GPU on the whole slide (r2h).  Equality with `WsiInstanceStore` on random tile sequences: tests/test_oracle_wsi.py."""

    def __init__(self) -> None:
        self._chunks: list[dict] = []          # kept instances per tile, already in WSI space
        self._starts: list[int] = []
        self._boxes = np.empty((4096, 4), dtype=np.int64)
        self._alive = np.zeros(4096, dtype=bool)
        self._chunk_boxes = np.empty((256, 4), dtype=np.int64)     # bounding box of every chunk's instances (WSI space)
        self.n = 0

    def __len__(self) -> int:
        return int(self._alive[: self.n].sum())

    def merge_tile(self, cols: dict, *, tile_bounds, tile_flag, tile_mode: int, margin, consume: bool = False) -> tuple[int, int]:
        """`consume=True`: the caller hands the tile's arrays over — the vertex array is shifted to WSI space in place instead of into a
        fresh copy (a slide's ~180 MB of vertices would otherwise be written a second time, into newly faulted pages)."""
        n = len(cols["box"])
        if n == 0:    # :2903-2904 — an empty cross tile evicts nothing either
            return 0, 0
        tile_bounds = np.asarray(tile_bounds)
        tile_tl = tile_bounds[:2].astype(np.int64)
        tile_shape = tile_bounds[2:] - tile_bounds[:2]
        drop = select_tile_removals(cols["box"], tile_shape, tile_flag, tile_mode, margin)
        evicted = 0
        if tile_mode == 3 and self._chunks:  # noqa: PLR2004
            # a cross tile evicts the accumulated instances its inner margin lines cut.  Only tiles placed near it can hold such
            # instances: the lines are tested against the bounding box of every earlier tile's kept set first (one vectorised test over
            # ~2 000 chunk boxes), then against the instances of the few chunks that qualify — not against all 680 k accumulated boxes
            # per cross tile (r2n: 5.8 of the replay's 8.2 s).
            lines = margin_lines(margin, tile_shape, tile_tl)
            cb = self._chunk_boxes[: len(self._chunks)]
            near = np.zeros(len(cb), dtype=bool)
            for line in lines:
                near |= boxes_intersect(cb, line)
            for c in np.nonzero(near)[0].tolist():
                start, kk = self._starts[c], len(self._chunks[c]["box"])
                hit = np.zeros(kk, dtype=bool)
                for line in lines:
                    hit |= boxes_intersect(self._boxes[start: start + kk], line)
                hit &= self._alive[start: start + kk]
                evicted += int(hit.sum())
                self._alive[start: start + kk][hit] = False
        keep = np.nonzero(~drop)[0]
        k = len(keep)
        if k == 0:
            return 0, evicted
        off = cols["cont_off"]
        xy = cols["cont_xy"]
        # the tile's vertex array is kept whole (shifted to WSI space in one pass); the kept instances address it through (lo, hi) —
        # no gather of the surviving vertices (a slide has ~20 M of them)
        if k == n:
            box, cen, prob, typ, lo, hi = cols["box"], cols["centroid"], cols["prob"], cols["type"], off[:-1], off[1:]
        else:
            box, cen, prob, typ, lo, hi = (cols["box"][keep], cols["centroid"][keep], cols["prob"][keep], cols["type"][keep],
                                           off[:-1][keep], off[1:][keep])
        if consume and xy.flags.writeable:
            xy_wsi = np.add(xy, tile_tl.astype(xy.dtype), out=xy)
        else:
            xy_wsi = xy + tile_tl.astype(xy.dtype)
        chunk = {"box": box + np.concatenate([tile_tl, tile_tl]), "centroid": cen + tile_tl.astype(np.float64), "prob": prob, "type": typ,
                 "cont_xy": xy_wsi, "cont_lo": lo, "cont_hi": hi}
        while self.n + k > len(self._boxes):
            self._boxes = np.concatenate([self._boxes, np.empty_like(self._boxes)])
            self._alive = np.concatenate([self._alive, np.zeros_like(self._alive)])
        self._boxes[self.n: self.n + k] = chunk["box"]
        self._alive[self.n: self.n + k] = True
        if len(self._chunks) == len(self._chunk_boxes):
            self._chunk_boxes = np.concatenate([self._chunk_boxes, np.empty_like(self._chunk_boxes)])
        b = chunk["box"]
        self._chunk_boxes[len(self._chunks), :2] = b[:, :2].min(0)
        self._chunk_boxes[len(self._chunks), 2:] = b[:, 2:].max(0)
        self._chunks.append(chunk)
        self._starts.append(self.n)
        self.n += k
        return k, evicted

    def append_all(self, cols: dict) -> None:
        """`_process_full_wsi` (:999-1076): one post-processing call over the whole map, nothing to merge or shift."""
        self.merge_tile(cols, tile_bounds=np.array([0, 0, 1 << 40, 1 << 40]), tile_flag=[0, 0, 0, 0], tile_mode=0, margin=0)

    def to_info_dict(self, offset=(0, 0)) -> dict:
        """{key: object array over the surviving instances, insertion order} like `WsiInstanceStore.to_info_dict`."""
        alive = self._alive[: self.n]
        total = int(alive.sum())
        if total == 0:
            return {}
        dx, dy = (int(v) for v in offset)
        # one pass over the chunks to lay the survivors out as whole-slide columns, then every object column is filled at C speed from
        # row views / slices (np.fromiter) — no per-chunk np.split, whose 700 k `swapaxes` calls were most of this function's time
        sels = [np.nonzero(alive[start: start + len(ch["box"])])[0] for start, ch in zip(self._starts, self._chunks)]
        box = self._boxes[: self.n][alive] + np.array([dx, dy, dx, dy])       # `_boxes` is the chunks' boxes back to back
        cen = np.concatenate([ch["centroid"] for ch in self._chunks])[alive] + np.array([dx, dy], dtype=np.float64)
        prob = np.concatenate([ch["prob"] for ch in self._chunks])[alive]
        typ = np.concatenate([ch["type"] for ch in self._chunks])[alive]

        def contour_views():       # slices of every chunk's own vertex array: the 20 M vertices of a slide are never copied
            for ch, sel in zip(self._chunks, sels):
                xy = ch["cont_xy"]
                if dx or dy:
                    xy = (xy + np.array([dx, dy], dtype=xy.dtype)).astype(xy.dtype)
                yield from map(xy.__getitem__, map(slice, ch["cont_lo"][sel].tolist(), ch["cont_hi"][sel].tolist()))

        out = {
            "box": np.fromiter(iter(box), dtype=object, count=total),
            "centroid": np.fromiter(iter(cen), dtype=object, count=total),
            "contours": np.fromiter(contour_views(), dtype=object, count=total),
        }
        typed = typ >= 0
        probs = np.array(prob.tolist(), dtype=object)
        types = np.array(typ.tolist(), dtype=object)
        probs[~typed] = None
        types[~typed] = None
        out["prob"], out["type"] = probs, types
        return {key: out[key] for key in INFO_KEYS}


def tile_instance_dict(info_dict: dict) -> dict:
    """`_get_inst_info_dicts` (:3061-3083) for one task: {i + 1: {key: values[i]}} from a tile's packed `info_dict` columns."""
    keys = list(info_dict.keys())
    n = len(info_dict[keys[0]]) if keys else 0
    return {i + 1: {key: info_dict[key][i] for key in keys} for i in range(n)}


def update_tile_based_predictions_array(new_pred: np.ndarray, wsi_pred: np.ndarray, bounds, offset=(0, 0),
                                        max_inst_value: int | None = None, seg_type: str = "instance") -> int | None:
    """`_update_tile_based_predictions_array` (:3160-3212) for one task: paste a tile's prediction map into the WSI array in place.
    Instance maps keep what is already there, skip every new instance that overlaps it, and shift the new ids by the running
    `max_inst_value` — an order-dependent recurrence, hence replayed in tile order on one rank."""
    x_start, y_start, x_end, y_end = (int(v) for v in bounds)
    x_start, x_end = x_start + int(offset[0]), x_end + int(offset[0])
    y_start, y_end = y_start + int(offset[1]), y_end + int(offset[1])
    max_h, max_w = wsi_pred.shape
    x_end, y_end = min(x_end, max_w), min(y_end, max_h)
    new = new_pred[0: y_end - y_start, 0: x_end - x_start]
    if seg_type == "instance":
        prev = wsi_pred[y_start:y_end, x_start:x_end]
        overlap = (new > 0) & (prev > 0)
        max_inst_value = 0 if max_inst_value is None else max_inst_value
        keep = new > 0
        if np.any(overlap):
            ids = np.unique(new[overlap])
            ids = ids[ids > 0]
            keep = (new > 0) & (~np.isin(new, ids))
        combined = prev.copy()
        combined[keep] = new[keep] + max_inst_value
        new = combined
        max_inst_value = (np.max(new) + max_inst_value) if np.any(keep) else max_inst_value
    wsi_pred[y_start:y_end, x_start:x_end] = new
    return max_inst_value


PASTE_SCRATCH = 4096   # scratch slots at the end of `paste_instance_tile_tensor`'s table


def paste_instance_tile_tensor(new_pred, wsi_pred, bounds, max_inst_value, lut):
    """`update_tile_based_predictions_array` (instance maps) on torch tensors — the same order-dependent recurrence, branch-free so that
    it needs no host synchronisation per tile: on the B200 the gathered tile maps stay on the device, the WSI map is pasted there
    (2 209 tiles of a 20 000-patch slide: ~1 s instead of ~11 s of numpy `unique` / `isin` / masked copies on the host, r2j) and comes
    back once.  int32 arithmetic wraps exactly like the reference's numpy int32 scalars do.

    new_pred (h, w) int32, wsi_pred (H, W) int32 [updated in place], max_inst_value: 0-dim int32 tensor, lut: int32 scratch of at least
    new_pred.max() + 1 + PASTE_SCRATCH entries (zero on entry, zero on return).  Returns the new max_inst_value tensor."""
    import torch

    x_start, y_start, x_end, y_end = (int(v) for v in bounds)
    max_h, max_w = wsi_pred.shape
    x_end, y_end = min(x_end, max_w), min(y_end, max_h)
    new = new_pred[0: y_end - y_start, 0: x_end - x_start]
    prev = wsi_pred[y_start:y_end, x_start:x_end]
    pos = new > 0
    overlap = pos & (prev > 0)
    ids = new.to(torch.int64)
    # ids of new instances that touch something already pasted: lut[id] = 1 (id 0 never qualifies: overlap implies new > 0)
    # plain scatter (all writes to an index carry the same value: 1 for a touched id, 0 to slot 0 from every other pixel) — an
    # accumulating index_put would sort a million indices per tile (r2k: 40 ms per tile)
    # (pixels without overlap write a 0 into one of 4096 scratch slots at the END of the table, spread by pixel position: a million
    # writes to one address serialise — r2l: 4 ms per tile)
    n_ids = lut.numel() - PASTE_SCRATCH
    flat = torch.arange(ids.numel(), device=ids.device, dtype=torch.int64).reshape(ids.shape)
    target = torch.where(overlap, ids, n_ids + (flat & (PASTE_SCRATCH - 1)))
    lut.scatter_(0, target.reshape(-1), overlap.reshape(-1).to(lut.dtype))
    keep = pos & (lut[ids] == 0)
    combined = torch.where(keep, new + max_inst_value, prev)
    new_max = torch.where(keep.any(), combined.max() + max_inst_value, max_inst_value)
    lut.zero_()
    wsi_pred[y_start:y_end, x_start:x_end] = combined
    return new_max
