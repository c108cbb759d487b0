"""ctypes binding of libtiab200.so (C ABI declared in include/tiab200.h).

The library is built in-tree by ``__graft_entry__.build()``.  There is NO CPU fallback:
if the shared object is missing, or a compute entry point is called without a GPU, the
call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libtiab200.so"

TIAB_OK = 0
STAIN_OK, STAIN_EMPTY_TISSUE_MASK, STAIN_DEGENERATE, STAIN_UNUSED_ATOM = 0, 1, 2, 3
METHOD_ID = {"macenko": 0, "vahadane": 1, "ruifrok": 2, "custom": 3}


class StainParams(C.Structure):
    _fields_ = [
        ("method", C.c_int),
        ("y_threshold", C.c_int),
        ("angular_percentile", C.c_double),
        ("regularizer", C.c_double),
        ("fixed_matrix", C.c_double * 6),
    ]


MAX_TAPS = 49


class ConvDesc(C.Structure):
    """Mirror of tiab_conv_desc_t (include/tiab200.h)."""

    _fields_ = [
        ("x", C.c_void_p), ("x_rows", C.c_longlong), ("x_ld", C.c_longlong), ("cin", C.c_int), ("x_lo_off", C.c_int),
        ("w", C.c_void_p), ("w_ld", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("taps", C.c_int), ("tap_off", C.c_int * MAX_TAPS),
        ("Hin", C.c_int), ("Win", C.c_int), ("y0", C.c_int), ("y1", C.c_int), ("x0", C.c_int), ("x1", C.c_int),
        ("Hd", C.c_int), ("Wd", C.c_int), ("dy", C.c_int), ("dx", C.c_int),
        ("Hr", C.c_int), ("Wr", C.c_int), ("ry", C.c_int), ("rx", C.c_int),
        ("out16", C.c_void_p), ("ld16", C.c_int), ("coff16", C.c_int), ("lo16", C.c_int),
        ("out32", C.c_void_p), ("ld32", C.c_int), ("coff32", C.c_int),
        ("out16b", C.c_void_p), ("ld16b", C.c_int), ("coff16b", C.c_int), ("lo16b", C.c_int),
        ("scale2", C.c_void_p), ("shift2", C.c_void_p), ("bias", C.c_void_p),
        ("res32", C.c_void_p), ("res16", C.c_void_p), ("ldres", C.c_int), ("coffres", C.c_int), ("reslo", C.c_int),
        ("relu", C.c_int),
        ("pre_scale", C.c_void_p), ("pre_shift", C.c_void_p), ("pre_relu", C.c_int), ("x_is_f32", C.c_int), ("split", C.c_int),
        ("groups", C.c_int), ("group_rows", C.c_int), ("x_grouped", C.c_int),
        ("tap_coff", C.c_int * MAX_TAPS), ("s2d_c", C.c_int),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_scale", C.c_void_p), ("head_shift", C.c_void_p),
        ("head_out", C.c_void_p * 4), ("head_nout", C.c_int * 4), ("head_ld", C.c_int * 4),
        ("out_scale", C.c_void_p),
    ]


class TiabError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()
_tls = threading.local()

_u8p = C.POINTER(C.c_uint8)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/tiab200.h declares
SIGNATURES = {
    "tiab_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "tiab_destroy": (None, [_vp]),
    "tiab_last_error": (C.c_char_p, [_vp]),
    "tiab_version": (C.c_char_p, []),
    "tiab_stain_fit": (C.c_int, [_vp, C.POINTER(StainParams), _vp, C.c_int, C.c_int, _dp, _dp, _ip, _vp]),
    "tiab_stain_transform_batch": (
        C.c_int,
        [_vp, C.POINTER(StainParams), _vp, _vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _vp, _vp, _vp, _vp, _vp],
    ),
    "tiab_stain_transform_host": (
        C.c_int,
        [_vp, C.POINTER(StainParams), _vp, _vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _vp, _vp, _vp],
    ),
    "tiab_stain_fit_host": (C.c_int, [_vp, C.POINTER(StainParams), _vp, C.c_int, C.c_int, _dp, _dp, _ip]),
    "tiab_stain_tissue_mask": (C.c_int, [_vp, C.POINTER(StainParams), _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "tiab_stain_concentrations": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _dp, _vp, _vp]),
    "tiab_stain_augment": (C.c_int, [_vp, _vp, _vp, C.c_longlong, _dp, _dp, _dp, _vp, _vp]),
    "tiab_reinhard_fit": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _dp, _dp, _vp]),
    "tiab_reinhard_transform_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _vp, _vp, _vp]),
    "tiab_reinhard_set_path": (C.c_int, [_vp, C.c_int]),
    "tiab_lab_convert": (C.c_int, [_vp, _vp, _vp, C.c_longlong, C.c_int, _vp]),
    "tiab_dbg_lab_convert": (C.c_int, [_vp, _vp, C.c_longlong, C.c_int]),
    "tiab_dbg_reinhard_maps": (C.c_int, [_vp, C.c_double, _dp, _dp, _dp, _vp]),
    "tiab_conv_gemm": (C.c_int, [_vp, C.POINTER(ConvDesc), _vp]),
    "tiab_im2col": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int] + [C.c_int] * 12 + [_vp, C.c_int, C.c_int, C.c_int, _vp]),
    "tiab_bn_relu": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_longlong, C.c_int, _vp, C.c_int,
                               C.c_int, C.c_int, _vp]),
    "tiab_upsample_add": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp] + [C.c_int] * 6 + [C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]
                          + [C.c_int] * 6 + [_vp]),
    "tiab_maxpool3x3s2": (C.c_int, [_vp, _vp] + [C.c_int] * 6 + [_vp] + [C.c_int] * 6 + [_vp]),
    "tiab_pointwise_head": (C.c_int, [_vp, _vp, C.c_longlong, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp,
                                      C.c_int, C.c_int, _vp]),
    "tiab_gap_fc_softmax": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "tiab_hovernet_heads": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_longlong, _vp, _vp, _vp, _vp]),
    "tiab_softmax_up2_crop": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "tiab_net_create": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(_vp)]),
    "tiab_net_info": (C.c_int, [_vp, _ip, C.c_int]),
    "tiab_net_infer": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "tiab_net_destroy": (None, [_vp]),
    "tiab_merge_batch_to_canvas": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "tiab_stitch_patches": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "tiab_hovernet_postproc": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tiab_hovernet_postproc_scaled": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp, C.c_int, _vp, _vp, _vp,
                                                _vp, _vp, _vp, _vp]),
    "tiab_hovernet_contours": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "tiab_hovernet_set_watershed": (C.c_int, [_vp, C.c_int]),
    "tiab_hovernet_get_watershed": (C.c_int, [_vp]),
    "tiab_dbg_watershed": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "tiab_dbg_watershed_local": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "tiab_dbg_watershed_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "tiab_dbg_sobel_kernels": (C.c_int, [_vp, _vp]),
    "tiab_dbg_sobel_kernels_k": (C.c_int, [C.c_int, _vp, _vp]),
    "tiab_dbg_trace_contour": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int]),
    "tiab_dbg_postproc_fetch": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "tiab_stain_set_path": (C.c_int, [_vp, C.c_int]),
    "tiab_last_launch_count": (C.c_int, [_vp]),
    "tiab_dbg_percentile_u8": (C.c_int, [_vp, C.c_double, _dp]),
    "tiab_dbg_contrast_lut": (C.c_int, [_vp, C.c_uint, _vp]),
    "tiab_dbg_eigh3": (C.c_int, [_dp, _dp, _dp]),
    "tiab_dbg_lasso2": (C.c_int, [_dp, _dp, C.c_double, C.c_double, _dp]),
    "tiab_dbg_macenko_from_angles": (C.c_int, [_dp, C.c_double, C.c_double, _dp]),
    "tiab_dbg_tables": (C.c_int, [_vp, _vp]),
}


def lib() -> C.CDLL:
    """Load libtiab200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not LIB_PATH.exists():
                    raise TiabError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                        " (there is no CPU fallback)"
                    )
                dll = C.CDLL(str(LIB_PATH))
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(dll, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = dll
    return _lib


def handle(device: int = 0) -> int:
    """Per-thread, per-device library handle (the handle is not thread-safe)."""
    cache = getattr(_tls, "handles", None)
    if cache is None:
        cache = _tls.handles = {}
    if device not in cache:
        h = _vp()
        rc = lib().tiab_create(int(device), C.byref(h))
        if rc != TIAB_OK or not h:
            raise TiabError(f"tiab_create(device={device}) failed (rc={rc}): no usable CUDA device — there is no CPU fallback")
        cache[device] = h
    return cache[device]


def check(h, rc: int) -> None:
    if rc != TIAB_OK:
        msg = lib().tiab_last_error(h)
        raise TiabError(f"libtiab200 error {rc}: {msg.decode() if msg else '?'}")


def dptr(arr: np.ndarray, typ=_dp):
    return arr.ctypes.data_as(typ)
