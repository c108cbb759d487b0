"""Shared scenario builders for the WSI tile-mode tests (host logic; no GPU)."""
from __future__ import annotations

import copy
from types import SimpleNamespace

import numpy as np


def random_tile_records(rng: np.random.Generator, tile_shape, margin: int, n: int) -> list[dict]:
    """Instance records in tile space; a third of the boxes are snapped onto the tile border / margin coordinates so that the closed
    interval comparisons (touching, exactly-inside) are exercised, not only generic positions."""
    w, h = int(tile_shape[0]), int(tile_shape[1])
    snaps_x = [0, 1, margin - 1, margin, margin + 1, w - margin - 1, w - margin, w - margin + 1, w - 1, w]
    snaps_y = [0, 1, margin - 1, margin, margin + 1, h - margin - 1, h - margin, h - margin + 1, h - 1, h]
    recs = []
    for k in range(n):
        bw, bh = int(rng.integers(3, 48)), int(rng.integers(3, 48))
        x0 = int(rng.integers(0, max(w - bw, 1)))
        y0 = int(rng.integers(0, max(h - bh, 1)))
        mode = int(rng.integers(0, 6))
        if mode == 0:
            x0 = min(max(int(rng.choice(snaps_x)), 0), max(w - bw, 0))
        elif mode == 1:
            x0 = min(max(int(rng.choice(snaps_x)) - bw, 0), max(w - bw, 0))
        elif mode == 2:
            y0 = min(max(int(rng.choice(snaps_y)), 0), max(h - bh, 0))
        elif mode == 3:
            y0 = min(max(int(rng.choice(snaps_y)) - bh, 0), max(h - bh, 0))
        box = np.array([x0, y0, x0 + bw, y0 + bh], dtype=np.int64)
        cont = np.array([[x0, y0], [x0 + bw - 1, y0], [x0 + bw - 1, y0 + bh - 1], [x0, y0 + bh - 1]], dtype=np.int32)
        recs.append({"box": box, "centroid": np.array([x0 + bw / 2.0, y0 + bh / 2.0]), "contours": cont,
                     "prob": float(rng.random()), "type": int(rng.integers(0, 6))})
    return recs


def pack_columns(recs: list[dict]) -> dict:
    """A tile's `info_dict` as HoVerNet.postproc packs it: one object column per key (empty float arrays when there is no instance)."""
    if not recs:
        e = np.empty(shape=0)
        return {"box": e, "centroid": e, "contours": e, "prob": e, "type": e}
    out = {}
    for key in ("box", "centroid", "contours", "prob", "type"):
        col = np.empty(len(recs), dtype=object)
        for i, r in enumerate(recs):
            col[i] = r[key]
        out[key] = col
    return out


def scenario(seed: int, image_shape=(2500, 2150), tile_shape=(1024, 1024), patch_output_shape=(164, 164), margin=128, density=14,
             tile_info=None):
    """[(bounds, flag, mode, records)] in processing order for a seeded random slide."""
    from oracle import wsi_oracle as wo

    rng = np.random.default_rng(seed)
    sets = tile_info if tile_info is not None else wo.get_tile_info(image_shape, tile_shape, patch_output_shape, margin)
    tiles = []
    for bounds, flag, mode in wo.build_tile_tasks(sets):
        shape = (int(bounds[2] - bounds[0]), int(bounds[3] - bounds[1]))
        n = int(rng.integers(0, density)) if rng.random() > 0.1 else 0
        tiles.append((np.asarray(bounds), np.asarray(flag), int(mode), random_tile_records(rng, shape, margin or 0, n)))
    return tiles


def replay_reference(m, tiles, margin):
    """Drive the UNMODIFIED reference functions (oracle/ref_shim.load_wsi_merge) exactly like `_process_tile_mode` :1230-1279."""
    ioconfig = SimpleNamespace(margin=margin)
    acc: dict = {}
    for bounds, flag, mode, recs in tiles:
        post = ({"task_type": "nuclei_segmentation", "info_dict": pack_columns(copy.deepcopy(recs))},)
        inst_dict = m._get_inst_info_dicts(post)[0]
        new, rem = m._compute_info_dict_for_merge(inst_dict=inst_dict, tile_mode=mode, ref_inst_info_dict=acc, ioconfig=ioconfig,
                                                  tile_shape=bounds[2:] - bounds[:2], tile_tl=bounds[:2], tile_flag=flag)
        acc.update(new)
        for uid in rem:
            acc.pop(uid, None)
    return list(acc.values())


def records_equal(a: list[dict], b: list[dict]) -> bool:
    if len(a) != len(b):
        return False
    for ra, rb in zip(a, b):
        for key in ("box", "centroid", "contours"):
            if not np.array_equal(np.asarray(ra[key]), np.asarray(rb[key])):
                return False
        if ra["prob"] != rb["prob"] or ra["type"] != rb["type"]:
            return False
    return True
