"""The work-list floods on the device — version 3 (lane-0 8-ary heap in shared memory) and version 4 (warp-cooperative 32-ary heap,
the default since r2) — and the global-memory versions 1 / 2: all bit-exact against the oracle and each other on 164-px maps, a 512-px
tile and both scale factors; versions 3 / 4 additionally on 300 adversarial inputs saturated with exact ties (plateaus, quantised values,
multi-pixel markers, markers outside the mask) and the hand-built known answers, run on the GPU through tiab_dbg_watershed_device.
Prints the timings (`-s`)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import hovernet_postproc_oracle as po
from tiatoolbox_b200.synthetic import hovernet_head_maps

pytestmark = pytest.mark.gpu


def _dev_flood(lib, impl, dist, markers, mask):
    from tiatoolbox_b200 import _lib

    d = np.ascontiguousarray(dist, dtype=np.float64)
    m = np.ascontiguousarray(markers, dtype=np.int32)
    k = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(d.shape, np.int32)
    hd = _lib.handle(0)
    _lib.check(hd, lib.tiab_dbg_watershed_device(hd, d.ctypes.data, m.ctypes.data, k.ctypes.data, d.shape[0], d.shape[1], impl, out.ctypes.data))
    return out


@pytest.mark.parametrize("impl", [3, 4])
def test_device_flood_with_exact_ties_and_known_answers(built_lib, impl):
    rng = np.random.default_rng(3)
    for trial in range(300):
        h, w = int(rng.integers(4, 40)), int(rng.integers(4, 40))
        levels = int(rng.choice([1, 2, 3, 8]))                     # 1 level = one big plateau: every comparison is a tie on the value
        dist = rng.integers(0, levels, (h, w)).astype(np.float64) * 0.25 - 1.0
        mask = rng.random((h, w)) < rng.uniform(0.5, 1.0)
        markers = np.zeros((h, w), np.int32)
        for lab in range(1, int(rng.integers(2, 9))):               # multi-pixel markers (all age 0), some outside the mask
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            markers[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 4))] = lab
        want = po.watershed_c(dist, markers, mask)
        got = _dev_flood(built_lib, impl, dist, markers, mask)
        assert np.array_equal(got, want), (trial, h, w, levels)
    # signed zeros and mixed signs must order like IEEE doubles (the integer image of the key canonicalises -0.0)
    dist = np.array([[0.0, -0.0, 0.5, -0.5, -0.0, 0.0, -1.0, 1e-300, -1e-300]])
    markers = np.array([[1, 0, 0, 0, 0, 0, 0, 0, 2]])
    assert np.array_equal(_dev_flood(built_lib, impl, dist, markers, np.ones_like(markers)), po.watershed(dist, markers, np.ones_like(markers)))
    img = np.array([[0.0, 1.0, 2.0, 5.0, 3.0, 1.5, 0.5]])
    markers = np.array([[1, 0, 0, 0, 0, 0, 2]])
    assert _dev_flood(built_lib, impl, img, markers, np.ones_like(markers)).tolist() == [[1, 1, 1, 1, 2, 2, 2]]
    img = np.zeros((1, 6))
    markers = np.array([[1, 0, 0, 0, 0, 2]])
    assert _dev_flood(built_lib, impl, img, markers, np.ones((1, 6))).tolist() == [[1, 1, 1, 2, 2, 2]]
    img = np.zeros((3, 5))
    markers = np.zeros((3, 5), int)
    markers[1, 0], markers[0, 4] = 7, 9
    mask = np.ones((3, 5), bool)
    mask[:, 2] = False
    mask[0, 4] = False
    out = _dev_flood(built_lib, impl, img, markers, mask)
    assert (out[:, :2] == 7).all() and (out[:, 2:] == 0).all()
    # a large random landscape with a few markers: thousands of queue operations in one work item (deep heaps)
    big = np.round(rng.normal(size=(150, 170)) * 4) / 4
    mk = np.zeros(big.shape, np.int32)
    for lab in range(1, 12):
        y, x = int(rng.integers(0, 148)), int(rng.integers(0, 168))
        mk[y:y + 2, x:x + 2] = lab
    msk = rng.random(big.shape) < 0.93
    assert np.array_equal(_dev_flood(built_lib, impl, big, mk, msk), po.watershed_c(big, mk, msk))


def test_all_floods_bit_exact_and_timed(built_lib):
    import torch

    from tiatoolbox_b200.models import postproc as pp

    def run(maps, scale=1.0):
        inst, table = pp.proc_np_hv_batch(*[torch.from_numpy(a).cuda() for a in maps], scale_factor=scale)
        return inst.cpu().numpy(), {k: v.cpu().numpy() for k, v in table.items()}

    def timed(maps):
        d = [torch.from_numpy(a).cuda() for a in maps]
        for _ in range(2):
            pp.proc_np_hv_batch(*d)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            pp.proc_np_hv_batch(*d)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 5

    def rows_equal(ta, tb):   # rows past count[n] are never written
        assert np.array_equal(ta["count"], tb["count"])
        for k in ta:
            if k == "count":
                continue
            for n, c in enumerate(ta["count"]):
                assert np.array_equal(ta[k][n, :c], tb[k][n, :c]), (k, n)

    small = hovernet_head_maps(16, 164, seed=0)
    big = hovernet_head_maps(1, 512, seed=3)
    res, times = {}, {}
    default_impl = pp.get_watershed_impl()
    try:
        for impl in (1, 2, 3, 4):
            pp.set_watershed_impl(impl)
            res[impl] = (run(small), run(big)[0], run(small, 0.5)[0])
            times[impl] = timed(small)
    finally:
        pp.set_watershed_impl(default_impl)
    print("\npost-processing of 16 maps (164 px, ~40 nuclei): " + ", ".join(f"flood v{i} {t:.2f} ms" for i, t in times.items()))
    for i in range(16):
        assert np.array_equal(res[1][0][0][i], po.proc_np_hv(small[0][i], small[1][i])), i
    assert np.array_equal(res[1][1][0], po.proc_np_hv(big[0][0], big[1][0]))
    for impl in (2, 3, 4):
        assert np.array_equal(res[impl][0][0], res[1][0][0]), impl
        assert np.array_equal(res[impl][1], res[1][1]), impl
        assert np.array_equal(res[impl][2], res[1][2]), impl
        rows_equal(res[1][0][1], res[impl][0][1])
