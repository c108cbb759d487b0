"""Patch-mode NucleusInstanceSegmentor pipeline logic without a GPU: double-buffered result copies, the capped instance-table / contour-vertex
transfer with its full-fetch fallback for crowded images, device-contour slicing, result packing, D2H byte accounting.  The CUDA pieces
(network, post-processing kernels, contour kernel, events, pinned memory) are replaced by CPU stand-ins built from the oracle and the
host-compiled device code; the engine code under test is the product's own `infer_patches` / `post_process_patches`."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import hovernet_postproc_oracle as po
from tests.test_wsi_engine_cpu import table_from_inst
from tiatoolbox_b200.models import postproc as pp
from tiatoolbox_b200.models.engine.nucleus_instance_segmentor import NucleusInstanceSegmentor
from tiatoolbox_b200.synthetic import hovernet_head_maps


class _Event:
    def record(self, *_a):
        pass

    def synchronize(self):
        pass


def _cpu_postproc(np_map, hv_map, tp_map=None, max_inst=None, want_table=True, scale_factor=1.0):  # noqa: ARG001
    n, h, w = np_map.shape[:3]
    max_inst = max(16, (h * w) // 10 + 1)
    insts, rows = [], []
    for i in range(n):
        inst = po.proc_np_hv(np_map[i].numpy(), hv_map[i].numpy(), scale_factor=scale_factor).astype(np.int32)
        insts.append(inst)
        rows.append(table_from_inst(inst, tp_map[i].numpy(), max_inst))
    return torch.from_numpy(np.stack(insts)), {k: torch.from_numpy(np.stack([r[k] for r in rows])) for k in rows[0]}


def _cpu_trace_contours(lib):
    def trace(inst, table, capacity=None):
        n, h, w = inst.shape
        max_inst = table["id"].shape[1]
        capacity = capacity or max(64, (h * w) // 2)
        off = np.zeros((n, max_inst), np.int32)
        ln = np.zeros((n, max_inst), np.int32)
        tot = np.zeros(n, np.int32)
        xy = np.zeros((n, capacity, 2), np.int32)
        maps = inst.numpy()
        for i in range(n):
            lab = np.ascontiguousarray(maps[i], dtype=np.int32)
            for k in range(int(table["count"][i])):
                buf = np.zeros(4 * lab.size, np.int32)
                cnt = lib.tiab_dbg_trace_contour(lab.ctypes.data, h, w, int(table["id"][i, k]), buf.ctypes.data, len(buf) // 2)
                off[i, k], ln[i, k] = tot[i], cnt
                if tot[i] + cnt <= capacity:
                    xy[i, tot[i]: tot[i] + cnt] = buf[: 2 * cnt].reshape(-1, 2)
                tot[i] += cnt
        table.update(cont_off=torch.from_numpy(off), cont_len=torch.from_numpy(ln), cont_total=torch.from_numpy(tot), cont_xy=torch.from_numpy(xy))
        return table
    return trace


class _CpuEngine(NucleusInstanceSegmentor):
    def __init__(self, heads, batch_size):  # noqa: D107 — bypasses EngineABC.__init__ (CUDA only)
        self.batch_size = batch_size
        self.device = torch.device("cpu")
        self.return_probabilities = False
        self._heads = heads
        self._cursor = 0
        outer = self

        class _Model:
            tasks = ["nuclei_segmentation"]

            @staticmethod
            def infer_batch_device(model, x, *, device):  # noqa: ARG004
                n = x.shape[0]
                lo = outer._cursor
                outer._cursor += n
                return tuple(torch.from_numpy(h[lo: lo + n]) for h in outer._heads)

        self.model = _Model()

    def _batches(self, patches):
        for i in range(0, len(patches), self.batch_size):
            yield torch.from_numpy(patches[i: i + self.batch_size])

    def _pinned(self, key, shape, dtype):
        cache = self.__dict__.setdefault("_pin_cache", {})
        k = (key, tuple(shape), dtype)
        if k not in cache:
            cache[k] = torch.empty(tuple(shape), dtype=dtype)
        return cache[k]


@pytest.mark.parametrize(("rows_fast", "xy_fast", "device_contours"), [(512, 4096, True), (8, 4096, True), (512, 64, True), (512, 4096, False)])
def test_patch_pipeline_on_cpu_stand_ins(built_lib, monkeypatch, rows_fast, xy_fast, device_contours):
    heads = hovernet_head_maps(5, 164, seed=33)
    monkeypatch.setattr(pp, "proc_np_hv_batch", _cpu_postproc)
    monkeypatch.setattr(pp, "trace_contours", _cpu_trace_contours(built_lib))
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *_a, **_k: SimpleNamespace(cuda_stream=0))
    eng = _CpuEngine(heads, batch_size=2)
    eng.table_rows_fast, eng.contour_vertices_fast, eng.device_contours = rows_fast, xy_fast, device_contours
    tiles = np.zeros((5, 256, 256, 3), np.uint8)
    raw = eng.infer_patches(tiles)
    out = eng.post_process_patches(raw, 5)
    assert out["predictions"].shape == (5, 164, 164) and out["predictions"].dtype == np.int32
    for i in range(5):
        ref_inst, ref_info = po.postproc(heads[0][i], heads[1][i], heads[2][i])
        assert np.array_equal(out["predictions"][i], ref_inst)
        assert len(out["box"][i]) == len(ref_info) > 8
        for j, rec in enumerate(ref_info.values()):
            assert np.array_equal(out["box"][i][j], rec["box"]) and np.array_equal(out["contours"][i][j], rec["contours"])
            assert out["contours"][i][j].dtype == np.int32 and out["type"][i][j] == rec["type"]
            np.testing.assert_allclose(out["centroid"][i][j], rec["centroid"], atol=1e-9)
    # the byte counter reflects the capped copies (+ the full table when a fallback was needed)
    full_rows = 164 * 164 // 10 + 1
    assert eng.d2h_bytes > 5 * 164 * 164 * 4
    if rows_fast == 8 or xy_fast == 64:
        assert eng.d2h_bytes > 5 * full_rows * 16          # at least the full box columns came across
    else:
        # three batches of two fixed-size records: count / cont_total, f64 columns, instance map, int32 columns, contour vertices
        assert eng.d2h_bytes == 6 * 4 * (2 + 2 * 512 * 3 + 164 * 164 + 512 * 8 + 2 * 4096)


# ---- world_size 2 (gloo): contiguous patch shards, all-gather of the instance maps + object gather of the records ----------------------
def _worker(rank: int, world: int, port: int, q):
    import os

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tiatoolbox_b200 import _lib

    heads = hovernet_head_maps(5, 164, seed=33)
    pp.proc_np_hv_batch = _cpu_postproc
    pp.trace_contours = _cpu_trace_contours(_lib.lib())
    torch.cuda.Event = _Event
    torch.cuda.current_stream = lambda *_a, **_k: SimpleNamespace(cuda_stream=0)
    lo, hi = NucleusInstanceSegmentor.shard_bounds(5, rank, world)
    eng = _CpuEngine(tuple(h[lo:hi] for h in heads), batch_size=2)      # the stand-in network serves this rank's shard in order
    eng.labels = None
    out = eng.run(np.zeros((5, 256, 256, 3), np.uint8))
    q.put((rank, lo, hi, out["predictions"], [[np.asarray(b) for b in boxes] for boxes in out["box"]], int(eng.d2h_bytes)))
    dist.barrier()
    dist.destroy_process_group()


def test_patch_pipeline_world2(built_lib):
    import os

    import torch.multiprocessing as mp

    heads = hovernet_head_maps(5, 164, seed=33)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted((r[1], r[2]) for r in res) == [(0, 3), (3, 5)]
    for _, _, _, preds, boxes, d2h in res:            # every rank returns the complete result
        assert preds.shape == (5, 164, 164) and d2h > 0
        for i in range(5):
            ref_inst, ref_info = po.postproc(heads[0][i], heads[1][i], heads[2][i])
            assert np.array_equal(preds[i], ref_inst) and len(boxes[i]) == len(ref_info)
            assert all(np.array_equal(a, r["box"]) for a, r in zip(boxes[i], ref_info.values()))


def test_result_pool_returns_a_block_only_when_every_view_is_gone():
    """`predictions` sits on recycled host memory (engine_abc.HostResultPool): a block may be reused only after the result array AND
    all views derived from it (slices, reshapes, torch.from_numpy) have been collected; the free list is bounded."""
    import gc

    import torch

    from tiatoolbox_b200.models.engine.engine_abc import HostResultPool

    pool = HostResultPool(max_free_bytes=3 * 8192)
    a = pool.take((8, 16, 16), np.int32)
    a[:] = 7
    row, flat, tt = a[3], a.reshape(-1)[5:9], torch.from_numpy(a)
    del a
    gc.collect()
    assert not pool.free
    del row, flat
    gc.collect()
    assert not pool.free
    assert int(tt.sum()) == 7 * 8 * 16 * 16
    del tt
    gc.collect()
    assert [n for n, _ in pool.free] == [8192]
    b = pool.take((8, 16, 16), np.int32)       # same pages again
    assert not pool.free and b.flags.writeable and b.shape == (8, 16, 16)
    c = pool.take((8, 16, 16), np.int32)       # pool empty -> a fresh block, not the live one
    b[:] = 1
    c[:] = 2
    assert int(b.max()) == 1 and int(c.min()) == 2
    del b, c
    gc.collect()
    d = pool.take((7, 16, 16), np.int32)       # a slightly smaller request reuses a block (<= 25 % slack) ...
    assert len(pool.free) == 1 and d.shape == (7, 16, 16) and d.nbytes == 7168
    e = pool.take((2, 16, 16), np.int32)       # ... a much smaller one does not
    assert len(pool.free) == 1
    del d, e
    gc.collect()
    assert pool.free_bytes() <= pool.max_free_bytes
    held = [pool.take((8, 16, 16), np.int32) for _ in range(6)]
    del held
    gc.collect()
    assert pool.free_bytes() <= pool.max_free_bytes and len(pool.free) == 3      # bounded: oldest blocks are dropped
