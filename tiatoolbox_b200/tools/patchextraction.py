"""Patch tiling coordinates (mirror of tiatoolbox/tools/patchextraction.py:488-613, `PatchExtractor.get_coordinates`).

Pure index arithmetic on the host: the WSI-mode engines use it for the patch grid (outputs on a stride grid covering
ceil(dim / stride) cells, x fastest; the input box is the output box expanded by (in - out) // 2) and, with
patch = stride = tile, for the post-processing tile grid (multi_task_segmentor.py:1412-1417).  Same argument names,
(width, height) order, validation and error messages as the reference."""
from __future__ import annotations

import numpy as np


def _invalid_shape(shape: np.ndarray) -> bool:
    """patchextraction.py `validate_shape`: True when the value is NOT a 2-vector of non-negative integers."""
    return (not np.issubdtype(shape.dtype, np.integer)) or np.size(shape) > 2 or np.any(shape < 0)  # noqa: PLR2004


class PatchExtractor:
    """Only the static coordinate helpers of the reference class live on the hot path (reading pixels is WSIReader file I/O)."""

    @staticmethod
    def get_coordinates(patch_output_shape=None, image_shape=None, patch_input_shape=None, stride_shape=None, *,
                        input_within_bound: bool = False, output_within_bound: bool = False):
        """-> input boxes (N, 4) [start_x, start_y, end_x, end_y] (and output boxes when `patch_output_shape` is given)."""
        return_output_bound = patch_output_shape is not None
        image_shape_arr = np.array(image_shape)
        patch_input_shape_arr = np.array(patch_input_shape)
        if patch_output_shape is None:
            output_within_bound = False
            patch_output_shape_arr = patch_input_shape_arr
        else:
            patch_output_shape_arr = np.array(patch_output_shape)
        stride_shape_arr = np.array(stride_shape)
        for name, arr in (("image_shape", image_shape_arr), ("patch_input_shape", patch_input_shape_arr),
                          ("patch_output_shape", patch_output_shape_arr), ("stride_shape", stride_shape_arr)):
            if _invalid_shape(arr):
                msg = f"Invalid `{name}` value {arr}."
                raise ValueError(msg)
        if np.any(patch_input_shape_arr < patch_output_shape_arr):
            msg = (f"`patch_input_shape` must larger than `patch_output_shape` "
                   f"{patch_input_shape_arr} must > {patch_output_shape_arr}.")
            raise ValueError(msg)
        if np.any(stride_shape_arr < 1):
            msg = f"`stride_shape` value {stride_shape_arr} must > 1."
            raise ValueError(msg)
        # ceil(dim / stride) cells per axis, x fastest (np.meshgrid(x, y) flattened row-major)
        nx = int(np.ceil(image_shape_arr[0] / stride_shape_arr[0]))
        ny = int(np.ceil(image_shape_arr[1] / stride_shape_arr[1]))
        xs = np.arange(nx, dtype=np.int64) * int(stride_shape_arr[0])
        ys = np.arange(ny, dtype=np.int64) * int(stride_shape_arr[1])
        output_tl = np.stack([np.tile(xs, ny), np.repeat(ys, nx)], axis=-1)
        output_br = output_tl + patch_output_shape_arr[None]
        input_tl = output_tl - ((patch_input_shape_arr - patch_output_shape_arr) // 2)[None]
        input_br = input_tl + patch_input_shape_arr[None]
        drop = np.zeros(len(output_tl), dtype=bool)
        if output_within_bound:
            drop |= np.any(output_br > image_shape_arr[None], axis=1)
        if input_within_bound:
            drop |= np.any(input_br > image_shape_arr[None], axis=1)
            drop |= np.any(input_tl < 0, axis=1)
        input_bounds = np.concatenate([input_tl[~drop], input_br[~drop]], axis=-1)
        if return_output_bound:
            return input_bounds, np.concatenate([output_tl[~drop], output_br[~drop]], axis=-1)
        return input_bounds
