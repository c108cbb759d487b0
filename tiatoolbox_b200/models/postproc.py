"""HoVer-Net post-processing through libtiab200.so (device) + the small host-side tail.

`proc_np_hv_batch` = HoVerNet._proc_np_hv for a batch (hovernet.py:502-616) entirely on the GPU;
`instance_info` assembles get_instance_info's dictionary (hovernet.py:618-748): box / centroid /
type / prob come from the device table; contours from `trace_contours` (device border following, csrc/contour_math.cuh) when the
table carries them, else from cv2.findContours on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from tiatoolbox_b200 import _lib


def proc_np_hv_batch(np_map: torch.Tensor, hv_map: torch.Tensor, tp_map: torch.Tensor | None = None, max_inst: int | None = None,
                     want_table: bool = True, scale_factor: float = 1.0):
    """np_map (N,H,W,1), hv_map (N,H,W,2), tp_map (N,H,W,1) float32 CUDA tensors ->
    inst_map (N,H,W) int32 CUDA tensor and (optionally) the per-image instance table as CUDA tensors."""
    assert np_map.is_cuda and np_map.dtype == torch.float32
    n, h_, w_ = np_map.shape[:3]
    dev = np_map.device
    np_map, hv_map = np_map.contiguous(), hv_map.contiguous()
    tp_map = None if tp_map is None else tp_map.contiguous()
    inst = torch.empty((n, h_, w_), dtype=torch.int32, device=dev)
    if max_inst is None:
        # every kept marker has >= ceil(10 * scale_factor^2) pixels (hovernet.py:565, 613)
        max_inst = max(16, (h_ * w_) // max(int(np.ceil(10 * scale_factor * scale_factor)), 1) + 1)
    table = None
    if want_table:
        table = {
            "count": torch.empty((n,), dtype=torch.int32, device=dev),
            "id": torch.empty((n, max_inst), dtype=torch.int32, device=dev),
            "box": torch.empty((n, max_inst, 4), dtype=torch.int32, device=dev),
            "centroid": torch.empty((n, max_inst, 2), dtype=torch.float64, device=dev),
            "type": torch.empty((n, max_inst), dtype=torch.int32, device=dev),
            "prob": torch.empty((n, max_inst), dtype=torch.float64, device=dev),
        }
    hd = _lib.handle(dev.index or 0)
    rc = _lib.lib().tiab_hovernet_postproc_scaled(
        hd, np_map.data_ptr(), hv_map.data_ptr(), None if tp_map is None else tp_map.data_ptr(), n, h_, w_, float(scale_factor),
        inst.data_ptr(), max_inst,
        None if table is None else table["count"].data_ptr(), None if table is None else table["id"].data_ptr(),
        None if table is None else table["box"].data_ptr(), None if table is None else table["centroid"].data_ptr(),
        None if table is None else table["type"].data_ptr(), None if table is None else table["prob"].data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(hd, rc)
    return inst, table


def set_watershed_impl(impl: int, device: int = 0) -> None:
    """Flood implementation used by `proc_np_hv_batch` on this thread's handle of `device`: 1 = binary heap, one thread per foreground
    component (the round-1 default); 2 = 8-ary heap with the total order (value, age, index), one warp per component; 3 = the same
    flood with each component's bounding box and heap staged in shared memory, persistent blocks over a work list; 4 = work list +
    shared memory with the whole warp driving a 32-ary heap (csrc/postproc.cu: pp_watershed_v4).  All four give identical labels
    (tests/test_watershed_v2_gpu.py)."""
    hd = _lib.handle(device)
    _lib.check(hd, _lib.lib().tiab_hovernet_set_watershed(hd, int(impl)))


def get_watershed_impl(device: int = 0) -> int:
    return int(_lib.lib().tiab_hovernet_get_watershed(_lib.handle(device)))


def trace_contours(inst: torch.Tensor, table: dict, capacity: int | None = None) -> dict:
    """Contours of every tabulated instance on the device (the `cv2.findContours(..., RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0]` of
    hovernet.py:685-693): adds `cont_off`, `cont_len` (N, max_inst), `cont_total` (N,) and `cont_xy` (N, capacity, 2) int32 tensors to
    `table` (vertices in image coordinates) and returns it.  `capacity` = vertices per image (default H*W // 2)."""
    n, h_, w_ = inst.shape
    dev = inst.device
    max_inst = table["id"].shape[1]
    capacity = int(capacity) if capacity is not None else max(64, (h_ * w_) // 2)
    table["cont_off"] = torch.empty((n, max_inst), dtype=torch.int32, device=dev)
    table["cont_len"] = torch.empty((n, max_inst), dtype=torch.int32, device=dev)
    table["cont_total"] = torch.empty((n,), dtype=torch.int32, device=dev)
    table["cont_xy"] = torch.empty((n, capacity, 2), dtype=torch.int32, device=dev)
    hd = _lib.handle(dev.index or 0)
    _lib.check(hd, _lib.lib().tiab_hovernet_contours(
        hd, inst.data_ptr(), n, h_, w_, max_inst, table["count"].data_ptr(), table["id"].data_ptr(), table["box"].data_ptr(), capacity,
        table["cont_off"].data_ptr(), table["cont_len"].data_ptr(), table["cont_total"].data_ptr(), table["cont_xy"].data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream))
    return table


def fetch_debug(which: int, shape, dtype, device=0) -> np.ndarray:
    out = np.empty(shape, dtype)
    hd = _lib.handle(device)
    _lib.check(hd, _lib.lib().tiab_dbg_postproc_fetch(hd, which, out.ctypes.data, torch.cuda.current_stream(device).cuda_stream))
    return out


def instance_info(inst_map: np.ndarray, table_row: dict, offset=(0, 0), with_contours: bool = True) -> dict:
    """get_instance_info's dict for ONE image from the device table (numpy views of one image's rows)."""
    offset = np.asarray(offset)
    count = int(table_row["count"])
    info = {}
    if count == 0:
        return info
    device_contours = with_contours and "cont_xy" in table_row
    if device_contours and int(table_row["cont_total"]) > len(table_row["cont_xy"]):
        device_contours = False   # the image needed more vertices than the device buffer holds: trace on the host instead
    # column-wise work first (one numpy call per column instead of one per nucleus), then the per-nucleus records
    ids = table_row["id"][:count].tolist()
    raw_boxes = table_row["box"][:count].astype(np.int64)
    boxes = raw_boxes + np.concatenate([offset, offset])[None]
    centroids = table_row["centroid"][:count] + offset[None]
    types = table_row["type"][:count].tolist()
    probs = table_row["prob"][:count].tolist()
    if device_contours:
        lens = table_row["cont_len"][:count].tolist()
        offs = table_row["cont_off"][:count].tolist()
        shifted = (table_row["cont_xy"][: int(table_row["cont_total"])] + offset[None]).astype(np.int32)
    elif with_contours:
        import cv2
    for k in range(count):
        inst_id = ids[k]
        rec = {"box": None, "centroid": None, "contours": None, "prob": None, "type": None}
        if device_contours:
            if lens[k] < 3:  # hovernet.py:696-701 (a squeezed 1- or 2-point contour has shape[0] < 3)  # noqa: PLR2004
                continue
            rec["contours"] = shifted[offs[k]: offs[k] + lens[k]]
        elif with_contours:
            b = raw_boxes[k]
            crop = (inst_map[b[1]:b[3], b[0]:b[2]] == inst_id).astype(np.uint8)
            contour = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0][0].astype(np.int32)
            contour = np.squeeze(contour)
            if contour.shape[0] < 3 or len(contour.shape) != 2:  # hovernet.py:696-701  # noqa: PLR2004
                continue
            rec["contours"] = (contour + (b[:2] + offset)[None]).astype(np.int32)   # hovernet.py:705 adds in place: stays int32
        rec["centroid"] = centroids[k]
        rec["box"] = boxes[k]
        if types[k] >= 0:
            rec["type"] = types[k]
            rec["prob"] = probs[k]
        info[inst_id] = rec
    return info


def instance_columns(inst_map: np.ndarray, table_row: dict, offset=(0, 0)) -> dict:
    """`instance_info` as COLUMNS (engine/wsi_merge.columns_from_records layout): the instances of ONE image straight from the device
    table — no Python work per nucleus.  Needs the device contours (`trace_contours`); an image whose vertex buffer overflowed, or a
    table without contours, goes through `instance_info` (host cv2) and is converted."""
    from tiatoolbox_b200.models.engine.wsi_merge import columns_from_records

    count = int(table_row["count"])
    if "cont_xy" not in table_row or int(table_row["cont_total"]) > len(table_row["cont_xy"]):
        return columns_from_records(instance_info(inst_map, table_row, offset).values())
    off2 = np.asarray(offset, dtype=np.int64)
    lens = table_row["cont_len"][:count].astype(np.int64)
    keep = np.nonzero(lens >= 3)[0]          # hovernet.py:696-701: a 1- or 2-point contour drops the instance  # noqa: PLR2004
    k = len(keep)
    new_off = np.zeros(k + 1, dtype=np.int64)
    new_off[1:] = np.cumsum(lens[keep])
    src = table_row["cont_off"][:count].astype(np.int64)[keep]
    idx = np.repeat(src - new_off[:-1], lens[keep]) + np.arange(int(new_off[-1]), dtype=np.int64)
    xy = (table_row["cont_xy"][idx] + off2.astype(np.int32)).astype(np.int32) if k else np.zeros((0, 2), np.int32)
    types = table_row["type"][:count][keep].astype(np.int64)
    return {
        "box": table_row["box"][:count][keep].astype(np.int64) + np.concatenate([off2, off2]),
        "centroid": table_row["centroid"][:count][keep] + off2.astype(np.float64),
        "prob": table_row["prob"][:count][keep].astype(np.float64),
        "type": np.where(types >= 0, types, -1),
        "cont_xy": xy, "cont_off": new_off,
    }
