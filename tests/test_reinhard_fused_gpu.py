"""The experimental fused Reinhard kernel (one CTA per image, image resident in shared memory) against the default three-launch path and
the reference outputs, byte for byte; prints both timings.  The fused kernel is an opt-in alternative (`set_reinhard_path(1)`): parity-green on the B200 since r2a (profiles/r2a_gpu_tests_experimental.log); not the default because the three-launch path is as fast (0.34 vs 0.38 ms)."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "reinhard_golden.npz")


def test_fused_path_equals_default(built_lib):
    import torch

    from oracle import reinhard_oracle as ro
    from oracle.make_golden_reinhard import inputs
    from tiatoolbox_b200.synthetic import he_tiles
    from tiatoolbox_b200.tools.stainnorm import get_normalizer, set_reinhard_path

    norm = get_normalizer("reinhard")
    norm.fit(inputs()[0])
    tiles = he_tiles(300, 256, 256, seed=4)                       # more images than SMs: the persistent loop wraps
    flat = np.full((1, 256, 256, 3), 180, np.uint8)
    batch = np.concatenate([tiles[:150], flat, tiles[150:]])
    x = torch.from_numpy(batch).cuda()

    def timed():
        for _ in range(2):
            norm.transform_batch(x, raise_on_error=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            norm.transform_batch(x, raise_on_error=False)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 5

    try:
        set_reinhard_path(0)
        ref, info0 = norm.transform_batch(x, return_info=True, raise_on_error=False)
        t0 = timed()
        set_reinhard_path(1)
        got, info1 = norm.transform_batch(x, return_info=True, raise_on_error=False)
        t1 = timed()
        small = norm.transform_batch(inputs()[1])                  # 96 x 128 tiles: eligible for the fused path too
        odd = norm.transform(inputs()[2])                          # 67 x 45: not a multiple of 16 bytes -> default path
    finally:
        set_reinhard_path(0)
    assert torch.equal(got, ref) and torch.equal(info0["status"], info1["status"]) and int(info1["status"][150]) == 1
    assert torch.equal(info0["source_means"], info1["source_means"]) and torch.equal(info0["source_stds"], info1["source_stds"])
    assert np.array_equal(got[150].cpu().numpy(), flat[0])
    for i in (0, 77, 299):
        assert np.array_equal(got[i if i < 150 else i + 1].cpu().numpy(), ro.transform(tiles[i], norm.target_means, norm.target_stds)), i
    assert np.array_equal(small, GOLD["tiles_out"]) and np.array_equal(odd, GOLD["odd_out"])
    print(f"\nReinhard, 301 tiles 256x256: three launches {t0:.3f} ms, fused {t1:.3f} ms")
