"""Generate tests/golden/stitch_golden.npz: outputs of the UNMODIFIED reference merge_batch_to_canvas
(tiatoolbox/models/engine/semantic_segmentor.py:1141-1183, loaded by oracle/ref_shim.load_semseg_merge) on seeded inputs.
Run in the build container (needs /root/reference):  python oracle/make_golden_stitch.py"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402


def cases():
    """(name, blocks, output_locations, merged_shape): regular overlapping rows, ragged right edge, zero blocks, 3-fold overlap."""
    rng = np.random.default_rng(1234)
    out = []
    # 1. stride = 3/4 patch: two-fold overlaps, last block clipped by the canvas width
    h, w, c, n = 8, 16, 2, 6
    blocks = rng.random((n, h, w, c), dtype=np.float32)
    xs = np.arange(n) * 12
    locs = np.stack([xs, np.zeros(n, int), xs + w, np.full(n, h)], axis=1)
    locs[:, 2] = np.minimum(locs[:, 2], 70)
    out.append(("overlap2", blocks, locs, (h, 70, c)))
    # 2. stride = 1/4 patch: four-fold overlaps, unsorted block order, some all-zero blocks (skipped, not counted), short last row
    n = 9
    blocks = rng.standard_normal((n, h, w, c)).astype(np.float32)
    blocks[[2, 5]] = 0
    xs = rng.permutation(n) * 4
    locs = np.stack([xs, np.zeros(n, int), xs + w, np.full(n, h)], axis=1)
    locs[-1, 3] = 5   # a block whose valid height is smaller (edge case of :1180)
    out.append(("overlap4_unsorted", blocks, locs, (h, 48, c)))
    # 3. single-channel float64, no overlap
    blocks = rng.random((3, 4, 4, 1))
    locs = np.array([[0, 0, 4, 4], [4, 0, 8, 4], [8, 0, 12, 4]])
    out.append(("disjoint_f64", blocks, locs, (4, 12, 1)))
    return out


def main():
    m = ref_shim.load_semseg_merge()
    data = {}
    for name, blocks, locs, shape in cases():
        canvas, count = m.merge_batch_to_canvas(blocks.copy(), locs.copy(), shape)
        data[f"{name}.blocks"], data[f"{name}.locs"], data[f"{name}.shape"] = blocks, locs, np.asarray(shape)
        data[f"{name}.canvas"], data[f"{name}.count"] = canvas, count
    out = ROOT / "tests" / "golden" / "stitch_golden.npz"
    np.savez_compressed(out, **data)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
