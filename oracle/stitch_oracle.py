"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's WSI stitching.

Follows tiatoolbox/models/engine/semantic_segmentor.py:
  merge_batch_to_canvas      :1141-1183  (restated literally; pinned bit-for-bit against the AST-extracted reference function and
                                          the reference's own known answers tests/engines/test_semantic_segmentor.py:283-333)
  merge_horizontal           :1186-1263  (dask containers replaced by Python lists of numpy row canvases)
  merge_vertical_chunkwise   :1398-1534  (dask blocks -> list items, zarr spill / progress bar dropped; the seam folding,
                                          count normalisation and clipping are restated line by line)
  clip_probabilities_to_shape :1537-1557

Parity status: `merge_batch_to_canvas` pinned; the horizontal / vertical drivers are "parity unpinned" against a running reference
(they need dask + zarr, which this image does not have) — they are checked by hand-built known answers and by the property
"equals the global overlap-add / count of all patches when at most two patch rows cover a pixel" (tests/test_oracle_stitch.py).
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

import numpy as np


def merge_batch_to_canvas(blocks: np.ndarray, output_locations: np.ndarray, merged_shape) -> tuple[np.ndarray, np.ndarray]:
    """semantic_segmentor.py:1141-1183.  Every block is placed at row 0 (one canvas per patch row); all-zero blocks are skipped
    and do NOT count (:1178); the count map is uint8 and wraps."""
    blocks = np.asarray(blocks)
    output_locations = np.asarray(output_locations)
    canvas = np.zeros(merged_shape, dtype=blocks.dtype)
    count = np.zeros((*merged_shape[:2], 1), dtype=np.uint8)
    for i, block in enumerate(blocks):
        xs, ys, xe, ye = (int(v) for v in output_locations[i])
        if not np.any(block):
            continue
        canvas[0: ye - ys, xs:xe, :] += block[0: ye - ys, 0: xe - xs, :]
        count[0: ye - ys, xs:xe, 0] += 1
    return canvas, count


def merge_horizontal(canvas_np: np.ndarray, output_locs: np.ndarray, change_indices, canvas_width: int | None = None):
    """semantic_segmentor.py:1186-1263 for a complete slide: returns (list of row canvases, list of row counts, output_locs_y)."""
    canvases, counts, locs_y = [], [], []
    output_locs = np.array(output_locs, dtype=np.int64, copy=True)
    start_idx = 0
    for c_idx in change_indices:
        output_locs_ = output_locs[: c_idx - start_idx]
        canvas_np_ = canvas_np[: c_idx - start_idx]
        batch_xs = np.min(output_locs_[:, 0], axis=0)
        merged_width = int(canvas_width) if canvas_width is not None else int(np.max(output_locs_[:, 2]) - batch_xs)
        output_locs_[:, 2] = np.minimum(output_locs_[:, 2], merged_width)
        merged_shape = (canvas_np_.shape[1], merged_width, canvas_np.shape[3])
        canvas_merge, count_merge = merge_batch_to_canvas(canvas_np_, output_locs_, merged_shape)
        canvases.append(canvas_merge)
        counts.append(count_merge)
        locs_y.append(output_locs_[:, (1, 3)].copy())
        canvas_np = canvas_np[c_idx - start_idx:]
        output_locs = output_locs[c_idx - start_idx:]
        start_idx = c_idx
    return canvases, counts, np.concatenate(locs_y) if locs_y else np.zeros((0, 2), np.int64)


def clip_probabilities_to_shape(probabilities: np.ndarray, output_shape, written_height: int):
    """semantic_segmentor.py:1537-1557."""
    if output_shape is None:
        return probabilities, written_height, False
    target_height, target_width = map(int, output_shape)
    remaining_height = target_height - written_height
    if remaining_height <= 0:
        return probabilities[:0], written_height, True
    clipped = probabilities[:remaining_height, :target_width, ...]
    if clipped.shape[0] == 0:
        return clipped, written_height, True
    return clipped, written_height + clipped.shape[0], False


def merge_vertical_chunkwise(canvases: list, counts: list, output_locs_y_: np.ndarray, output_shape=None) -> np.ndarray:
    """semantic_segmentor.py:1398-1534 with the row canvases held in lists: fold the seam rows of the next row canvas into the
    current one (:1462-1464), normalise by the count with 0 -> 1 (:1467-1468), clip (:1470-1476), continue with the rest of the
    next canvas (:1512-1513)."""
    y0s, y1s = np.unique(output_locs_y_[:, 0]), np.unique(output_locs_y_[:, 1])
    overlaps = np.append(y1s[:-1] - y0s[1:], 0)
    num_chunks = len(canvases)
    written_height = 0
    out = []
    curr_chunk = canvases[0].copy()
    curr_count = counts[0].copy()
    next_chunk = canvases[1].copy() if num_chunks > 1 else None
    next_count = counts[1].copy() if num_chunks > 1 else None
    for i, overlap in enumerate(overlaps):
        overlap = int(overlap)
        if next_chunk is not None and overlap > 0:
            curr_chunk[-overlap:] += next_chunk[:overlap]
            curr_count[-overlap:] += next_count[:overlap]
        curr_count = np.where(curr_count == 0, 1, curr_count)
        probabilities = curr_chunk / curr_count.astype(np.float32)
        probabilities, written_height, should_stop = clip_probabilities_to_shape(probabilities, output_shape, written_height)
        if should_stop:
            break
        out.append(probabilities)
        if next_chunk is not None:
            curr_chunk, curr_count = next_chunk[overlap:], next_count[overlap:]
        if i + 2 < num_chunks:
            next_chunk = canvases[i + 2].copy()
            next_count = counts[i + 2].copy()
        else:
            next_chunk, next_count = None, None
    return np.concatenate(out, axis=0) if out else np.zeros((0,), np.float32)


def stitch(blocks: np.ndarray, output_locs: np.ndarray, canvas_width: int | None = None, output_shape=None) -> np.ndarray:
    """The data flow of SemanticSegmentor.infer_wsi's merging (semantic_segmentor.py:560-631) for patches that are all in memory:
    patches arrive in row-major grid order; a new row canvas starts whenever ys changes."""
    output_locs = np.asarray(output_locs, dtype=np.int64)
    ys = output_locs[:, 1]
    change = list(np.nonzero(ys[1:] != ys[:-1])[0] + 1) + [len(ys)]
    canvases, counts, locs_y = merge_horizontal(np.asarray(blocks), output_locs, change, canvas_width)
    return merge_vertical_chunkwise(canvases, counts, locs_y, output_shape)
