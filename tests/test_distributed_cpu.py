"""world_size-2 gloo tests (CPU) of the engines' sharding + final gather logic."""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank: int, world: int, port: int, n_total: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tiatoolbox_b200.models.engine.engine_abc import EngineABC

    lo, hi = EngineABC.shard_bounds(n_total, rank, world)
    # per-patch "results": row i holds its global patch index
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None].repeat(1, 3, 2)
    gathered = EngineABC._all_gather_rows(local, n_total)
    objs = [None] * world
    dist.all_gather_object(objs, list(range(lo, hi)))
    q.put((rank, lo, hi, gathered.numpy(), objs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8, 1])
def test_shard_and_gather_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = sorted((r[1], r[2]) for r in res)
    assert bounds[0][0] == 0 and bounds[-1][1] == n_total and bounds[0][1] == bounds[1][0]  # contiguous, complete, disjoint
    for _, _, _, g, objs in res:
        assert g.shape == (n_total, 3, 2)
        assert np.array_equal(g[:, 0, 0], np.arange(n_total))  # row-major patch order preserved
        assert [i for part in objs for i in part] == list(range(n_total))


def test_shard_bounds_properties():
    from tiatoolbox_b200.models.engine.engine_abc import EngineABC

    for n in (0, 1, 5, 8, 20000):
        for world in (1, 2, 4, 8):
            spans = [EngineABC.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // world)
