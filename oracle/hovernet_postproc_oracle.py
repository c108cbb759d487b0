"""CPU oracle of HoVer-Net's post-processing (architecture/hovernet.py:502-748 of the reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two layers:
  * `remove_small_objects` and `watershed` restate the two scikit-image (>= 0.26, pinned in
    requirements/requirements.txt:29 of the reference, NOT installed / not vendored) functions that
    `_proc_np_hv` calls (hovernet.py:544, 614, 616).  scikit-image's algorithm is restated from its
    published implementation (skimage/morphology/misc.py remove_small_objects; skimage/segmentation/
    _watershed.py + _watershed_cy.pyx watershed_raveled: priority flood, 4-connectivity, heap key
    (image value, insertion age), pixels are labelled when they are PUSHED, markers outside the
    mask are dropped, neighbour order (-row, -1, +1, +row)).
    Parity status of these two: **unpinned** — every reference test that exercises them needs remote
    weights/images, so they are pinned only by hand-built known-answer cases and
    property tests (tests/test_oracle_postproc.py).  Exact float ties between different markers
    depend on scikit-image's heap internals and are treated as unspecified.
  * `proc_np_hv` / `get_instance_info` are a numpy + cv2 + scipy restatement of the rest of
    hovernet.py:536-748; when /root/reference is present the tests run the reference's own
    unmodified `_proc_np_hv` / `get_instance_info` (through oracle/ref_shim.py, with the two
    restated functions injected) against it.
"""
from __future__ import annotations

import heapq
import math

import cv2
import numpy as np
from scipy import ndimage


def remove_small_objects(label_img: np.ndarray, max_size: int) -> np.ndarray:
    """skimage.morphology.remove_small_objects(ar, max_size=...) for an integer label image:
    zero every label whose pixel count is <= max_size; no relabelling."""
    out = label_img.copy()
    sizes = np.bincount(label_img.ravel())
    too_small = sizes <= max_size
    out[too_small[label_img]] = 0
    return out


def watershed(image: np.ndarray, markers: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """skimage.segmentation.watershed(image, markers=markers, mask=mask) with the defaults
    (connectivity=1, compactness=0, watershed_line=False)."""
    img = np.asarray(image, dtype=np.float64)
    msk = np.asarray(mask).astype(bool)
    out = (np.asarray(markers) * msk).astype(np.int32)
    h, w = img.shape
    heap: list = []
    age = 0
    for idx in np.flatnonzero(out):
        heapq.heappush(heap, (img.flat[idx], 0, int(idx)))
    flat_out = out.ravel()
    flat_img = img.ravel()
    flat_msk = msk.ravel()
    while heap:
        _, _, p = heapq.heappop(heap)
        y, x = divmod(p, w)
        lab = flat_out[p]
        for q, ok in ((p - w, y > 0), (p - 1, x > 0), (p + 1, x < w - 1), (p + w, y < h - 1)):
            if not ok or not flat_msk[q] or flat_out[q]:
                continue
            age += 1
            flat_out[q] = lab
            heapq.heappush(heap, (flat_img[q], age, q))
    return out


# ---- the same two functions in plain C (oracle/csrc/postproc_oracle.c, built by __graft_entry__.build() with gcc into oracle/_build/) ----
_CLIB = None


def build_c(force: bool = False):
    """Compile oracle/csrc/postproc_oracle.c with gcc (a second, independently written restatement + compiled speed for the CPU arm)."""
    import subprocess
    from pathlib import Path

    here = Path(__file__).resolve().parent
    src, out = here / "csrc" / "postproc_oracle.c", here / "_build" / "libtiab_oracle.so"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    return out


def _clib():
    global _CLIB  # noqa: PLW0603
    if _CLIB is None:
        import ctypes as C

        lib = C.CDLL(str(build_c()))
        lib.tiab_oracle_watershed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.tiab_oracle_remove_small_objects.argtypes = [C.c_void_p, C.c_long, C.c_int]
        _CLIB = lib
    return _CLIB


def watershed_c(image: np.ndarray, markers: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """`watershed` through the C restatement (same semantics, same tie rule)."""
    img = np.ascontiguousarray(image, dtype=np.float64)
    mk = np.ascontiguousarray(markers, dtype=np.int32)
    msk = np.ascontiguousarray(np.asarray(mask).astype(bool).astype(np.uint8))
    out = np.empty(img.shape, np.int32)
    rc = _clib().tiab_oracle_watershed(img.ctypes.data, mk.ctypes.data, msk.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data)
    assert rc == 0
    return out


def remove_small_objects_c(label_img: np.ndarray, max_size: int) -> np.ndarray:
    out = np.ascontiguousarray(label_img, dtype=np.int32).copy()
    assert _clib().tiab_oracle_remove_small_objects(out.ctypes.data, out.size, int(max_size)) == 0
    return out.astype(label_img.dtype, copy=False)


def proc_np_hv(np_map: np.ndarray, hv_map: np.ndarray, *, debug: dict | None = None, scale_factor: float = 1.0) -> np.ndarray:
    """HoVerNet._proc_np_hv (hovernet.py:536-616); scale_factor 1 for HoVer-Net, 0.5 for HoVerNetPlus (hovernetplus.py:305)."""
    blb_raw = np_map[..., 0]
    h_dir_raw = hv_map[..., 0]
    v_dir_raw = hv_map[..., 1]
    blb = np.array(blb_raw >= 0.5, dtype=np.int32)
    blb = ndimage.label(blb)[0]
    blb = remove_small_objects(blb, max_size=9)
    blb[blb > 0] = 1
    h_dir = cv2.normalize(h_dir_raw, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    v_dir = cv2.normalize(v_dir_raw, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    ksize = int((20 * scale_factor) + 1)
    obj_size = math.ceil(10 * (scale_factor**2))
    sobel_h = cv2.Sobel(h_dir, cv2.CV_64F, 1, 0, ksize=ksize)
    sobel_v = cv2.Sobel(v_dir, cv2.CV_64F, 0, 1, ksize=ksize)
    sobel_h = 1 - cv2.normalize(sobel_h, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    sobel_v = 1 - cv2.normalize(sobel_v, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    overall = np.maximum(sobel_h, sobel_v)
    overall = overall - (1 - blb)
    overall[overall < 0] = 0
    dist = (1.0 - overall) * blb
    dist = -cv2.GaussianBlur(dist, (3, 3), 0)
    overall = np.array(overall >= 0.4, dtype=np.int32)
    marker = blb - overall
    marker[marker < 0] = 0
    marker = ndimage.binary_fill_holes(marker).astype("uint8")
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
    marker = cv2.morphologyEx(marker, cv2.MORPH_OPEN, kernel)
    marker = ndimage.label(marker)[0]
    marker = remove_small_objects(marker, max_size=obj_size - 1)
    if debug is not None:
        debug.update(blb=blb.copy(), dist=dist.copy(), marker=marker.copy(), h_dir=h_dir, v_dir=v_dir)
    return watershed(dist, markers=marker, mask=blb)


def get_bounding_box(img: np.ndarray) -> np.ndarray:
    """utils/misc.py:898-922: [start_x, start_y, end_x, end_y], end exclusive."""
    rows = np.any(img, axis=1)
    cols = np.any(img, axis=0)
    r_min, r_max = np.where(rows)[0][[0, -1]]
    c_min, c_max = np.where(cols)[0][[0, -1]]
    return np.array([c_min, r_min, c_max + 1, r_max + 1])


def get_instance_info(pred_inst: np.ndarray, pred_type: np.ndarray | None = None, offset=(0, 0)) -> dict:
    """HoVerNet.get_instance_info (hovernet.py:670-748)."""
    offset = np.asarray(offset)
    info = {}
    for inst_id in np.unique(pred_inst)[1:]:
        inst_map = pred_inst == inst_id
        box = get_bounding_box(inst_map)
        tl = box[:2] + offset
        crop = inst_map[box[1]:box[3], box[0]:box[2]].astype(np.uint8)
        mom = cv2.moments(crop)
        contour = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0][0].astype(np.int32)
        contour = np.squeeze(contour)
        if contour.shape[0] < 3 or len(contour.shape) != 2:
            continue
        centroid = np.array([mom["m10"] / mom["m00"], mom["m01"] / mom["m00"]]) + tl
        contour += tl[None]   # in place like hovernet.py:705 — the contour stays int32
        box = box.copy()
        box[:2] += offset
        box[2:] += offset
        info[int(inst_id)] = {"box": box, "centroid": centroid, "contours": contour, "prob": None, "type": None}
    if pred_type is not None:
        for inst_id, rec in info.items():
            b = rec["box"].copy()
            b[:2] -= offset
            b[2:] -= offset
            c0, r0, c1, r1 = b
            m = pred_inst[r0:r1, c0:c1] == inst_id
            types = pred_type[r0:r1, c0:c1][m]
            tl_, cnt = np.unique(types, return_counts=True)
            order = sorted(zip(tl_, cnt), key=lambda x: x[1], reverse=True)
            t = order[0][0]
            if t == 0 and len(order) > 1:
                t = order[1][0]
            rec["type"] = int(t)
            rec["prob"] = float(dict(order)[t] / (np.sum(m) + 1.0e-6))
    return info


# ---------------------------------------------------------------------------------------------------------------------
# HoVerNetPlus (architecture/hovernetplus.py): layer-segmentation clean-up, layer contours, two-task postproc
# ---------------------------------------------------------------------------------------------------------------------
def remove_small_objects_bool(mask: np.ndarray, max_size: int) -> np.ndarray:
    """skimage.morphology.remove_small_objects for a BOOLEAN array: connected components (1-connectivity, like ndimage.label's default)
    of at most max_size pixels are cleared."""
    lab = ndimage.label(mask)[0]
    sizes = np.bincount(lab.ravel())
    small = sizes <= max_size
    small[0] = False
    out = mask.copy()
    out[small[lab]] = False
    return out


def proc_ls(ls_map: np.ndarray) -> np.ndarray:
    """HoVerNetPlus._proc_ls (hovernetplus.py:135-194)."""
    ls_map = np.around(np.squeeze(ls_map)).astype("uint8")
    min_size, kernel_size = 20000, 20
    epith_all = np.where(ls_map >= 2, 1, 0).astype("uint8") > 0
    mask = np.where(ls_map >= 1, 1, 0).astype("uint8")
    epith_mask = remove_small_objects_bool(epith_all, max_size=min_size - 1).astype("uint8")
    epith_edited = (epith_mask * ls_map).astype("uint8")
    epith_edited_open = np.zeros_like(epith_edited).astype("uint8")
    k = np.ones((kernel_size, kernel_size))
    for i in [3, 2, 4]:
        tmp = np.where(epith_edited == i, 1, 0).astype("uint8")
        ep_open = cv2.morphologyEx(cv2.morphologyEx(tmp, cv2.MORPH_CLOSE, k), cv2.MORPH_OPEN, k)
        epith_edited_open[ep_open == 1] = i
    mask_open = cv2.morphologyEx(cv2.morphologyEx(mask, cv2.MORPH_CLOSE, k), cv2.MORPH_OPEN, k).astype("uint8")
    out = mask_open.copy()
    for i in range(2, 5):
        out[epith_edited_open == i] = i
    return out.astype("uint8")


def get_layer_info(pred_layer: np.ndarray, offset=(0, 0)) -> dict:
    """HoVerNetPlus._get_layer_info (hovernetplus.py:197-264)."""
    info, count = {}, 1
    for type_class in [v for v in np.unique(pred_layer) if v != 0]:
        layer = np.where(pred_layer == type_class, 1, 0).astype("uint8")
        bounding_box = get_bounding_box(layer)
        contours, _ = cv2.findContours(layer, cv2.RETR_TREE, cv2.CHAIN_APPROX_NONE)
        for cont in contours:
            sq = np.squeeze(cont)
            if sq.shape[0] < 3 or len(sq.shape) != 2:
                continue
            box = bounding_box.copy()
            box[:2] = box[:2] + offset
            box[2:] = box[2:] + offset
            info[count] = {"box": box, "contours": cont[:, 0, :] + offset, "type": type_class}
            count += 1
    return info


def postproc_plus(np_map, hv_map, tp_map, ls_map, offset=(0, 0)):
    """HoVerNetPlus.postproc (hovernetplus.py:267-410) -> (inst_map, nucleus info dict, layer map, layer info dict)."""
    inst = proc_np_hv(np_map, hv_map, scale_factor=0.5)
    layer = proc_ls(ls_map)
    pred_type = np.around(tp_map).astype("uint8")
    return inst, get_instance_info(inst, pred_type, offset), layer, get_layer_info(layer, offset)


def postproc(np_map, hv_map, tp_map=None, offset=(0, 0)):
    """HoVerNet.postproc (hovernet.py:751-859) -> (inst_map int32, info dict)."""
    pred_type = None if tp_map is None else np.around(tp_map).astype("uint8")
    inst = proc_np_hv(np_map, hv_map)
    return inst, get_instance_info(inst, pred_type, offset)
