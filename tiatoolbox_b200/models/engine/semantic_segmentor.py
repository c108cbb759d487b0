"""SemanticSegmentor (mirror of tiatoolbox/models/engine/semantic_segmentor.py): `fcn-tissue_mask`.

Patch mode as in the reference; in addition the reference's WSI-mode *merging* (semantic_segmentor.py:1141-1263, 1398-1534) runs on
the device: `merge_batch_to_canvas` / `stitch_patches` below keep the reference's function contract (numpy in -> numpy out, or CUDA
tensors in -> CUDA tensors out) and `SemanticSegmentor.run(image, patch_mode=False)` accepts an in-memory RGB region (the reference
reads regions through WSIReader, which is file I/O and outside the B200 path), tiles it like the reference's PatchExtractor grid,
infers the tiles and stitches the probability map without leaving the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from tiatoolbox_b200 import _lib
from tiatoolbox_b200.models.engine.engine_abc import EngineABC, cast_to_min_dtype


def _prep(blocks, output_locations):
    """-> (float32 contiguous CUDA tensor (N, h, w, C), int32 host array (N, 4), was_numpy)."""
    was_numpy = not isinstance(blocks, torch.Tensor)
    if was_numpy:
        arr = np.asarray(blocks)
        if arr.dtype != np.float32:
            msg = f"the device stitcher works on float32 patch predictions, got {arr.dtype}"
            raise TypeError(msg)
        t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    else:
        if not blocks.is_cuda or blocks.dtype != torch.float32:
            msg = "blocks must be a float32 CUDA tensor (or a float32 numpy array)"
            raise TypeError(msg)
        t = blocks.contiguous()
    if t.dim() != 4:
        msg = f"blocks must have shape (N, H, W, C), got {tuple(t.shape)}"
        raise ValueError(msg)
    locs = np.ascontiguousarray(np.asarray(output_locations).reshape(-1, 4).astype(np.int32))
    if len(locs) != t.shape[0]:
        msg = "one [start_x, start_y, end_x, end_y] row per block is required"
        raise ValueError(msg)
    return t, locs, was_numpy


def merge_batch_to_canvas(blocks, output_locations, merged_shape):
    """Device version of semantic_segmentor.py:1141-1183: merge one row of patch predictions (N, h, w, C) placed at
    [start_x, start_y, end_x, end_y] into a canvas `merged_shape` = (H, W, C) and a uint8 count map (H, W, 1); all-zero blocks are
    skipped and not counted.  float32 sums are formed in block order, i.e. bit-identical to the reference's `canvas[...] += block`."""
    t, locs, was_numpy = _prep(blocks, output_locations)
    H, W, Cc = (int(v) for v in merged_shape)
    n, bh, bw, c = t.shape
    if n and c != Cc:
        msg = f"merged_shape has {Cc} channels, blocks have {c}"
        raise ValueError(msg)
    dev = t.device
    canvas = torch.empty((H, W, Cc), dtype=torch.float32, device=dev)
    count = torch.empty((H, W, 1), dtype=torch.uint8, device=dev)
    h = _lib.handle(dev.index or 0)
    _lib.check(h, _lib.lib().tiab_merge_batch_to_canvas(
        h, t.data_ptr() if n else None, n, max(bh, 1), max(bw, 1), Cc, locs.ctypes.data_as(C.c_void_p) if n else None, H, W,
        canvas.data_ptr() if canvas.numel() else None, count.data_ptr() if count.numel() else None,
        torch.cuda.current_stream(dev).cuda_stream))
    if was_numpy:
        return canvas.cpu().numpy(), count.cpu().numpy()
    return canvas, count


def stitch_patches(blocks, output_locations, canvas_width: int | None = None, output_shape=None, out: torch.Tensor | None = None):
    """The merging of SemanticSegmentor.infer_wsi (merge_horizontal :1186-1263 per patch row, merge_vertical_chunkwise :1398-1534,
    clip_probabilities_to_shape :1537-1557) for patch predictions that are all in memory, in one device pass.

    blocks (N, h, w, C) float32 in row-major grid order (x fastest; a new patch row starts when start_y changes);
    returns the (H, W, C) float32 probability map = overlap sums / counts (uncovered pixels divide by 1).  `out`: an optional
    contiguous (H, W, C) float32 CUDA tensor to write into (e.g. a row slice of a WSI canvas) instead of a fresh allocation."""
    t, locs, was_numpy = _prep(blocks, output_locations)
    n, bh, bw, c = t.shape
    if n == 0:
        msg = "no patches to stitch"
        raise ValueError(msg)
    width = int(canvas_width) if canvas_width is not None else int(locs[:, 2].max())
    if output_shape is None:
        output_shape = (int(locs[:, 3].max() - locs[0, 1]), width)
    H, W = int(output_shape[0]), min(int(output_shape[1]), width)
    if out is None:
        out = torch.empty((H, W, c), dtype=torch.float32, device=t.device)
    elif (tuple(out.shape) != (H, W, c) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != t.device or was_numpy):
        msg = f"out must be a contiguous float32 tensor of shape {(H, W, c)} on the blocks' device"
        raise ValueError(msg)
    h = _lib.handle(t.device.index or 0)
    _lib.check(h, _lib.lib().tiab_stitch_patches(h, t.data_ptr(), n, bh, bw, c, locs.ctypes.data_as(C.c_void_p), width, H, W, out.data_ptr(),
                                                 torch.cuda.current_stream(t.device).cuda_stream))
    return out.cpu().numpy() if was_numpy else out


def patch_grid(image_shape, patch_shape, stride_shape) -> np.ndarray:
    """Top-left aligned row-major tile grid over an (H, W) region (x fastest, like tools/patchextraction.py:576-591 with
    within_bound=False): int64 (N, 4) [start_x, start_y, end_x, end_y]; the last row / column may reach past the region."""
    H, W = int(image_shape[0]), int(image_shape[1])
    ph, pw = int(patch_shape[0]), int(patch_shape[1])
    sy, sx = int(stride_shape[0]), int(stride_shape[1])
    ys = list(range(0, max(H - ph, 0) + 1, sy))
    xs = list(range(0, max(W - pw, 0) + 1, sx))
    if ys[-1] + ph < H:
        ys.append(ys[-1] + sy)
    if xs[-1] + pw < W:
        xs.append(xs[-1] + sx)
    return np.array([[x, y, x + pw, y + ph] for y in ys for x in xs], dtype=np.int64)


class SemanticSegmentor(EngineABC):
    """run(...) -> {"predictions": (N, h, w) argmax maps (bool for 2 classes), ["probabilities": (N, h, w, C)]}."""

    def infer_patches(self, patches: np.ndarray) -> torch.Tensor:
        outs = [type(self.model).infer_batch_device(self.model, x, device=self.device).clone() for x in self._batches(patches)]
        return torch.cat(outs) if outs else torch.empty((0, 1, 1, 1), device=self.device)

    def post_process_patches(self, raw: torch.Tensor, n_total: int) -> dict:
        # the argmax is taken on the device so that only 1 byte/pixel crosses the bus unless probabilities are asked for
        pred = self._all_gather_rows(raw.argmax(dim=-1).to(torch.uint8), n_total).cpu().numpy()
        out = {"predictions": cast_to_min_dtype(pred)}
        if self.return_probabilities:
            out["probabilities"] = self._all_gather_rows(raw, n_total).cpu().numpy()
        return out

    def run(self, images, *, patch_mode: bool = True, patch_input_shape=None, stride_shape=None, **kwargs):
        """Patch mode: EngineABC.run.  patch_mode=False with an in-memory (H, W, 3) uint8 region: tile -> infer -> device stitching
        (`run_region`); file / WSIReader inputs stay outside the B200 path."""
        if not patch_mode and isinstance(images, np.ndarray) and images.ndim == 3:
            if patch_input_shape is None:
                msg = "patch_input_shape=(h, w) is required to tile a region"
                raise ValueError(msg)
            for k, v in kwargs.items():
                if k not in ("masks", "input_resolutions", "ioconfig", "save_dir", "overwrite", "output_type"):
                    setattr(self, k, v)
            return self.run_region(images, patch_input_shape, stride_shape)
        return super().run(images, patch_mode=patch_mode, patch_input_shape=patch_input_shape, **kwargs)

    def run_region(self, image: np.ndarray, patch_input_shape, stride_shape=None, output_scale: float | None = None) -> dict:
        """WSI-mode data flow for an in-memory RGB region (H, W, 3) uint8: tile (edge tiles zero padded), infer, stitch on the device
        (`stitch_patches`), argmax.  `output_scale` = output pixels per input pixel of the model (probed from one tile if None)."""
        image = np.ascontiguousarray(image)
        if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
            msg = "run_region expects an (H, W, 3) uint8 array"
            raise ValueError(msg)
        ph, pw = int(patch_input_shape[0]), int(patch_input_shape[1])
        stride_shape = stride_shape or (ph // 2, pw // 2)
        locs_in = patch_grid(image.shape[:2], (ph, pw), stride_shape)
        Hp, Wp = int(locs_in[:, 3].max()), int(locs_in[:, 2].max())
        padded = np.zeros((Hp, Wp, 3), np.uint8)
        padded[: image.shape[0], : image.shape[1]] = image
        tiles = np.stack([padded[y0:y1, x0:x1] for x0, y0, x1, y1 in locs_in])
        self.model._plan_batch = self.batch_size
        probs = self.infer_patches(tiles)                                # (N, oh, ow, C) on the device
        oh, ow = int(probs.shape[1]), int(probs.shape[2])
        s = float(output_scale) if output_scale is not None else oh / ph
        if abs(oh - ph * s) > 1e-6 or abs(ow - pw * s) > 1e-6 or any(abs(v * s - round(v * s)) > 1e-6 for v in locs_in.ravel()):
            msg = "tile grid and model output scale do not line up (choose a stride that maps to whole output pixels)"
            raise ValueError(msg)
        locs_out = np.rint(locs_in * s).astype(np.int64)
        out_h, out_w = int(round(image.shape[0] * s)), int(round(image.shape[1] * s))
        prob = stitch_patches(probs.contiguous(), locs_out, canvas_width=int(locs_out[:, 2].max()), output_shape=(out_h, out_w))
        out = {"predictions": cast_to_min_dtype(prob.argmax(dim=-1).to(torch.uint8).cpu().numpy())}
        if self.return_probabilities:
            out["probabilities"] = prob.cpu().numpy()
        return out
