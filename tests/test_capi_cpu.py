"""CPU tests of the C-ABI library: it loads, exports every symbol include/tiab200.h declares, and
its host-callable scalar pieces (the same __host__ __device__ code the kernels' finalise steps run)
agree with the oracle.  No compute entry point is called (no GPU here)."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import stain_oracle as so

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported(built_lib):
    header = (ROOT / "include" / "tiab200.h").read_text()
    names = set(re.findall(r"\b(tiab_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 15
    from tiatoolbox_b200 import _lib

    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in tiab200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert built_lib.tiab_version() == b"0.1"


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and pointer-vs-scalar kind of every ctypes signature against the C prototype (a mismatch would only show as a crash on the GPU)."""
    import ctypes

    from tiatoolbox_b200 import _lib

    header = (ROOT / "include" / "tiab200.h").read_text()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(?:int|void|const char\s*\*)\s*(tiab_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 30
    for name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [q.strip() for q in params.split(",")]
        res, args = _lib.SIGNATURES[name]
        assert len(args) == len(plist), (name, len(args), plist)
        for a, q in zip(args, plist):
            is_ptr = "*" in q or "[" in q or q.startswith(("tiab_handle_t", "tiab_net_t"))
            ct_ptr = a in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(a, "_type_") and isinstance(a._type_, type)
            assert is_ptr == bool(ct_ptr), (name, q, a)
            if not is_ptr:
                want = {"int": ctypes.c_int, "int32_t": ctypes.c_int, "long long": ctypes.c_longlong, "double": ctypes.c_double,
                        "float": ctypes.c_float, "unsigned": ctypes.c_uint, "size_t": ctypes.c_size_t}
                base = re.sub(r"\bconst\b", "", q).strip().rsplit(" ", 1)[0].strip()
                assert a is want[base], (name, q, a)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/tiab200.h must compile as C99 (no torch / C++ types in the signatures) and as C++."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = str(ROOT / "include" / "tiab200.h")
    for args in (["-std=c99", "-x", "c"], ["-std=c++17", "-x", "c++"]):
        r = subprocess.run([gcc, *args, "-fsyntax-only", "-Wall", "-Werror", hdr], capture_output=True, text=True, check=False)
        assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tiatoolbox_b200 import _lib

    with pytest.raises(_lib.TiabError, match="no CPU fallback"):
        _lib.handle(0)


def test_product_never_imports_oracle():
    for p in (ROOT / "tiatoolbox_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p


def test_tables(built_lib):
    tabg = np.zeros(256, np.uint32)
    od = np.zeros(256, np.float32)
    assert built_lib.tiab_dbg_tables(tabg.ctypes.data, od.ctypes.data) == 0
    assert np.array_equal(tabg.astype(np.int64), so.TABG)
    ref = so.rgb2od(np.arange(256, dtype=np.uint8).reshape(1, -1, 1).repeat(3, 2).copy())[0, :, 0]
    assert np.allclose(od, ref, rtol=1e-6)


def test_percentile_and_contrast_lut(built_lib):
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(3, 5000)) * 3
        if trial % 4 == 0:
            vals = rng.integers(0, 256, n).astype(np.uint8)
        elif trial % 4 == 1:
            vals = np.clip(rng.normal(200, 30, n), 0, 255).astype(np.uint8)
        elif trial % 4 == 2:
            vals = np.full(n, int(rng.integers(0, 256)), np.uint8)
        else:
            vals = rng.choice(np.array([0, 255, 17], np.uint8), n)
        hist = np.bincount(vals, minlength=256).astype(np.uint32)
        for q in (2, 98, 50, 1, 99):
            out = C.c_double()
            assert built_lib.tiab_dbg_percentile_u8(hist.ctypes.data, q / 100.0, C.byref(out)) == 0
            assert out.value == np.percentile(vals, q), (trial, q)
        lut = np.zeros(256, np.uint8)
        built_lib.tiab_dbg_contrast_lut(hist.ctypes.data, n, lut.ctypes.data)
        img = vals.reshape(-1, 1, 3)
        assert np.array_equal(lut[img], so.contrast_enhancer(img, 2, 98))


def test_contrast_lut_reference_golden_vector(built_lib, stain_golden):
    img = stain_golden["contrast_in"]
    hist = np.bincount(img.ravel(), minlength=256).astype(np.uint32)
    lut = np.zeros(256, np.uint8)
    built_lib.tiab_dbg_contrast_lut(hist.ctypes.data, img.size, lut.ctypes.data)
    assert np.array_equal(lut[img], stain_golden["contrast_out"])


def test_eigh3(built_lib):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.normal(size=(50, 3)) * rng.uniform(0.01, 3, size=3)
        cov = np.cov(a, rowvar=False)
        ev = np.zeros(3)
        vec = np.zeros(9)
        built_lib.tiab_dbg_eigh3(cov.ravel().ctypes.data_as(C.POINTER(C.c_double)), ev.ctypes.data_as(C.POINTER(C.c_double)),
                                 vec.ctypes.data_as(C.POINTER(C.c_double)))
        w, v = np.linalg.eigh(cov)
        assert np.allclose(ev, w, rtol=1e-10, atol=1e-14)
        vec = vec.reshape(3, 3)
        for k in range(3):
            s = np.sign(vec[:, k] @ v[:, k])
            assert np.allclose(vec[:, k] * s, v[:, k], atol=1e-7)


def test_lasso2(built_lib):
    rng = np.random.default_rng(2)
    dp = C.POINTER(C.c_double)
    for t in range(500):
        n = int(10 ** rng.uniform(1, 6.5))
        d = rng.normal(size=(2, 50)) * rng.uniform(0.1, 3)
        if t % 3 == 0:
            d[1] = d[0] * rng.uniform(0.5, 1.5) + rng.normal(size=50) * 0.05
        y = rng.normal(size=50) * rng.uniform(0.01, 3)
        gram, c = np.ascontiguousarray(d @ d.T), np.ascontiguousarray(d @ y)
        out = np.zeros(2)
        built_lib.tiab_dbg_lasso2(gram.ctypes.data_as(dp), c.ctypes.data_as(dp), 0.1, float(n), out.ctypes.data_as(dp))
        assert np.allclose(out, so.lasso2(gram, c, 0.1, n), rtol=1e-9, atol=1e-12)


def test_macenko_from_angles(built_lib):
    rng = np.random.default_rng(3)
    dp = C.POINTER(C.c_double)
    for _ in range(50):
        e = np.linalg.qr(rng.normal(size=(3, 3)))[0][:, :2].copy()
        a, b = sorted(rng.uniform(-1.5, 1.5, 2))
        m = np.zeros(6)
        built_lib.tiab_dbg_macenko_from_angles(np.ascontiguousarray(e).ctypes.data_as(dp), a, b, m.ctypes.data_as(dp))
        v1 = e @ np.array([np.cos(a), np.sin(a)])
        v2 = e @ np.array([np.cos(b), np.sin(b)])
        he = np.array([v1, v2]) if v1[0] > v2[0] else np.array([v2, v1])
        he = he / np.linalg.norm(he, axis=1)[:, None]
        assert np.allclose(m.reshape(2, 3), he, atol=1e-12)
