"""WSI mode of MultiTaskSegmentor / NucleusInstanceSegmentor on the B200 (BASELINE config 5).

Reference data flow (tiatoolbox/models/engine/multi_task_segmentor.py): `infer_wsi` :477-731 reads patches through WSIReader (255-padded
outside the slide, dataset_abc.py:419-431), runs `infer_batch` per batch with a blocking D2H, and stitches WSI-scale head maps on
the host (dask / zarr spill); `post_process_wsi` :836-997 then slices those maps tile by tile (`_process_tile_mode` :1078-1287), runs
`postproc_func` per tile on a dask thread pool and merges the per-tile instance dictionaries in order.

Here the slide band of a rank is uploaded ONCE, and everything up to the per-tile instance tables stays in HBM:

  slide band (uint8, 255 border)  --strided view-->  patch batches  --stain normaliser / network-->  head outputs
      --device stitcher (csrc/stitch.cu)-->  head canvases (float32, H x W x {1, 2, 1})
      --tile slices-->  on-device watershed + instance table (csrc/postproc.cu)  --D2H-->  contours (cv2) + `wsi_merge` replay

Multi-GPU (one process per GPU): the post-processing tile grid is cut into contiguous bands of tile rows; a rank owns the grid tiles
of its band and the strips / cross-sections anchored in it, infers exactly the patch rows those tiles read (a halo of at most
2 x margin below the band is recomputed by both neighbours, no traffic during compute), and the ranks exchange only the final
per-tile results: one all-gather of the flattened int32 tile instance maps (NCCL) + an object all-gather of the ragged instance
records.  Every rank then replays the order-dependent merge (`wsi_merge.WsiInstanceStore`), which is deterministic, so the result is
identical for any number of ranks.

Inputs are in-memory RGB arrays: reading slide files (WSIReader / openslide / tifffile) is file I/O outside the B200 path.
"""
from __future__ import annotations

import numpy as np
import torch

from tiatoolbox_b200.models import postproc as pp
from tiatoolbox_b200.models.engine import wsi_merge as wm
from tiatoolbox_b200.tools.patchextraction import PatchExtractor


class WsiTilePlan:
    """Host-side geometry of one slide for one rank: patch grid, post-processing tiles, and what this rank owns."""

    def __init__(self, image_wh, ioconfig, rank: int = 0, world: int = 1) -> None:
        self.W, self.H = int(image_wh[0]), int(image_wh[1])
        self.margin = 0 if ioconfig.margin is None else int(ioconfig.margin)
        self.patch_in = tuple(int(v) for v in ioconfig.patch_input_shape)
        self.patch_out = tuple(int(v) for v in (ioconfig.patch_output_shape or ioconfig.patch_input_shape))
        self.stride = tuple(int(v) for v in ioconfig.stride_shape)
        self.in_boxes, self.out_boxes = PatchExtractor.get_coordinates(
            image_shape=(self.W, self.H), patch_input_shape=self.patch_in, patch_output_shape=self.patch_out, stride_shape=self.stride)
        tile_shape = ioconfig.tile_shape if ioconfig.tile_shape is not None else (self.W, self.H)
        # post_process_wsi :939-941: tile mode only when the map is larger than one tile; otherwise one "tile" = the whole map
        if self.H > int(tile_shape[0]) or self.W > int(tile_shape[1]):
            sets = wm.get_tile_info((self.W, self.H), tile_shape, self.patch_out, self.margin)
            self.tile_mode = True
        else:
            sets = [[np.array([[0, 0, self.W, self.H]], dtype=np.int64), np.zeros((1, 4), dtype=np.int32)]]
            self.tile_mode = False
        self.tasks = wm.build_tile_tasks(sets)
        # ---- ownership: bands of grid-tile rows; strips / crosses follow the tile row their top edge lies in
        grid = sets[0][0]
        row_ys = np.unique(grid[:, 1])
        tile_h = int(grid[0, 3] - grid[0, 1])
        per = -(-len(row_ys) // world)
        lo, hi = min(len(row_ys), rank * per), min(len(row_ys), (rank + 1) * per)
        rows_of = [int(np.clip(int(b[1]) // tile_h, 0, len(row_ys) - 1)) for b, _, _ in self.tasks]
        self.my_tasks = [i for i, r in enumerate(rows_of) if lo <= r < hi]
        # ---- the patch rows this rank must infer: every patch whose output rows meet the y-range its tiles read
        if self.my_tasks:
            ya = max(0, min(int(self.tasks[i][0][1]) for i in self.my_tasks))
            yb = min(self.H, max(int(self.tasks[i][0][3]) for i in self.my_tasks))
            sel = (self.out_boxes[:, 3] > ya) & (self.out_boxes[:, 1] < yb)
        else:
            sel = np.zeros(len(self.out_boxes), dtype=bool)
        self.my_patches = np.nonzero(sel)[0]
        if len(self.my_patches):
            self.canvas_y0 = int(self.out_boxes[self.my_patches, 1].min())
            self.canvas_y1 = min(self.H, int(self.out_boxes[self.my_patches, 3].max()))
            # image rows the band's patches read (may reach outside the slide: padded with 255)
            self.read_y0 = int(self.in_boxes[self.my_patches, 1].min())
            self.read_y1 = int(self.in_boxes[self.my_patches, 3].max())
            self.read_x0 = int(self.in_boxes[self.my_patches, 0].min())
            self.read_x1 = int(self.in_boxes[self.my_patches, 2].max())
        else:
            self.canvas_y0 = self.canvas_y1 = self.read_y0 = self.read_y1 = self.read_x0 = self.read_x1 = 0

    def patch_rows(self):
        """Yield (patch indices of one grid row) for this rank, top to bottom."""
        if not len(self.my_patches):
            return
        ys = self.out_boxes[self.my_patches, 1]
        for y in np.unique(ys):
            yield self.my_patches[ys == y]


class WsiTileModeMixin:
    """`run(image, patch_mode=False)` for instance-segmentation engines.  The three device primitives are methods so that the host
    logic (banding, stitching order, tile batching, gather, replay) can be exercised without a GPU by the gloo tests."""

    tile_batch = 16  # same-shape tiles post-processed per device call (the flood's serial critical path — the largest component — is shared by the batch)
    min_tile = 11    # smallest tile side tiab_hovernet_postproc accepts

    # ---- device primitives -------------------------------------------------------------------------------------------------
    def _wsi_forward(self, x: torch.Tensor) -> list[torch.Tensor]:
        """uint8 (n, h, w, 3) device patches -> head maps [(n, oh, ow, C) float32] (preproc_func = stain normaliser runs on the device)."""
        x = self._device_preproc(x)
        outs = type(self.model).infer_batch_device(self.model, x, device=self.device)
        return list(outs)

    def _wsi_stitch_row(self, blocks: torch.Tensor, locs: np.ndarray, width: int, out: torch.Tensor) -> None:
        """One patch row of head outputs -> rows of the head canvas (device overlap-add / count normalisation, csrc/stitch.cu)."""
        from tiatoolbox_b200.models.engine.semantic_segmentor import stitch_patches

        stitch_patches(blocks, locs, canvas_width=width, output_shape=(out.shape[0], out.shape[1]), out=out)

    def _wsi_postproc(self, np_map: torch.Tensor, hv_map: torch.Tensor, tp_map: torch.Tensor | None, scale_factor: float = 1.0):
        """(n, h, w, C) head tiles -> (inst (n, h, w) int32, instance table) on the device (csrc/postproc.cu)."""
        return pp.proc_np_hv_batch(np_map, hv_map, tp_map, scale_factor=scale_factor)

    # ---- stages ------------------------------------------------------------------------------------------------------------
    def _wsi_upload(self, image: np.ndarray, plan: WsiTilePlan) -> torch.Tensor:
        """The image rows / columns the band's patches read, as one device uint8 tensor with a 255 border where patches leave the
        slide (`read_bounds(..., pad_constant_values=255)`, dataset_abc.py:419-431)."""
        y0, y1, x0, x1 = plan.read_y0, plan.read_y1, plan.read_x0, plan.read_x1
        host = torch.full((y1 - y0, x1 - x0, 3), 255, dtype=torch.uint8)
        if self.device.type == "cuda":
            host = host.pin_memory()
        sy0, sy1, sx0, sx1 = max(y0, 0), min(y1, plan.H), max(x0, 0), min(x1, plan.W)
        host[sy0 - y0: sy1 - y0, sx0 - x0: sx1 - x0] = torch.from_numpy(image[sy0:sy1, sx0:sx1])
        return host.to(self.device, non_blocking=True)

    def _wsi_infer(self, band: torch.Tensor, plan: WsiTilePlan) -> list[torch.Tensor]:
        """All patch rows of the band -> head canvases [(canvas rows, W, C) float32] in HBM.

        Rows that do not overlap vertically (stride == patch_output: every registry ioconfig) are stitched row by row as they finish.
        With a vertical overlap (stride[1] < patch_output[1]) consecutive rows share output pixels that the reference's
        merge_vertical_chunkwise averages (semantic_segmentor.py:1398-1534): the head blocks of the whole band are kept and go through
        ONE `stitch_patches` call, which folds neighbouring rows exactly like the reference."""
        pw, ph = plan.patch_in
        Wp = band.shape[1]
        canvases: list[torch.Tensor] | None = None
        rowbufs: list[torch.Tensor] | None = None
        v_overlap = plan.stride[1] < plan.patch_out[1]
        kept_blocks: list[list[torch.Tensor]] = []
        kept_locs: list[np.ndarray] = []
        for idx in plan.patch_rows():
            nx = len(idx)
            ins, outs = plan.in_boxes[idx], plan.out_boxes[idx]
            # strided view of the row's (overlapping) input windows: no host loop, the copy happens in .contiguous() per batch
            yoff, xoff = int(ins[0, 1]) - plan.read_y0, int(ins[0, 0]) - plan.read_x0
            sx = plan.stride[0]
            base = band[yoff: yoff + ph]
            view = base.as_strided((nx, ph, pw, 3), (sx * 3, Wp * 3, 3, 1), base.storage_offset() + xoff * 3)
            if v_overlap:
                rowbufs = None                 # the row's blocks are kept until the band is stitched: fresh buffers per row
            for k0 in range(0, nx, self.batch_size):
                k1 = min(nx, k0 + self.batch_size)
                heads = self._wsi_forward(view[k0:k1].contiguous())
                if canvases is None:   # head count / channels are known after the first forward (infer_wsi :581-589)
                    ch = plan.canvas_y1 - plan.canvas_y0
                    need = sum(ch * plan.W * h.shape[3] * 4 for h in heads)
                    if self.device.type == "cuda":
                        free = torch.cuda.mem_get_info(self.device)[0]
                        if need > 0.9 * free:
                            msg = (f"the head canvases of this band need {need / 1e9:.1f} GB of HBM ({free / 1e9:.1f} GB free): run the slide "
                                   "on more ranks (bands shrink with the world size) or in regions; the B200 path keeps canvases resident "
                                   "instead of spilling to zarr like the reference")
                            raise MemoryError(msg)
                    canvases = [torch.zeros((ch, plan.W, h.shape[3]), dtype=torch.float32, device=self.device) for h in heads]
                if rowbufs is None or rowbufs[0].shape[0] < nx:
                    rowbufs = [torch.empty((nx, *h.shape[1:]), dtype=torch.float32, device=self.device) for h in heads]
                for buf, h in zip(rowbufs, heads):
                    buf[k0:k1].copy_(h)
            locs = outs.astype(np.int32)
            if v_overlap:
                kept_blocks.append([buf[:nx] for buf in rowbufs])
                kept_locs.append(locs)
                continue
            y0 = int(outs[0, 1]) - plan.canvas_y0
            y1 = min(int(outs[0, 3]) - plan.canvas_y0, canvases[0].shape[0])
            for buf, canvas in zip(rowbufs, canvases):
                self._wsi_stitch_row(buf[:nx], locs, plan.W, canvas[y0:y1])
        if v_overlap and kept_blocks:
            locs = np.concatenate(kept_locs)
            locs[:, [1, 3]] -= plan.canvas_y0          # band-local rows: the stitcher writes canvas rows [0, canvas height)
            for hidx, canvas in enumerate(canvases):
                self._wsi_stitch_row(torch.cat([blk[hidx] for blk in kept_blocks]), locs, plan.W, canvas)
        return canvases or []

    def _wsi_postproc_tiles(self, canvases: list[torch.Tensor], plan: WsiTilePlan, *, keep_maps: bool):
        """This rank's tiles -> [(task index, tile instance COLUMNS (wsi_merge.columns_from_records layout), tile inst map (device int32) | None, extra)] in task order; `extra` is
        None for HoVer-Net and (layer region dict, layer map uint8 | None) for the two-task HoVerNetPlus (hovernetplus.py:267-410 per tile)."""
        results = []
        if not canvases:
            return results
        has_tp = len(canvases) > 2
        two_task = getattr(self.model, "num_layers", None) is not None
        scale = 0.5 if two_task else 1.0

        def clipped(i):
            x0, y0, x1, y1 = (int(v) for v in plan.tasks[i][0])
            return max(x0, 0), max(y0, 0), min(x1, plan.W), min(y1, plan.H)

        group: list[int] = []

        def flush():
            if not group:
                return
            tiles = []
            for i in group:
                x0, y0, x1, y1 = clipped(i)
                tiles.append([c[y0 - plan.canvas_y0: y1 - plan.canvas_y0, x0:x1] for c in canvases])
            stacks = [torch.stack([t[h] for t in tiles]).contiguous() for h in range(len(canvases))]
            th, tw = stacks[0].shape[1:3]
            if th < self.min_tile or tw < self.min_tile:
                # a sliver left over at the slide's right / bottom edge: the device kernels (21-tap Sobel) need >= 11 px per side, so the
                # sliver is extended with background (all heads 0) and the instance map cropped back
                pad = (0, 0, 0, max(self.min_tile - tw, 0), 0, max(self.min_tile - th, 0))
                stacks = [torch.nn.functional.pad(s, pad).contiguous() for s in stacks]
            inst, table = self._wsi_postproc(stacks[0], stacks[1], stacks[2] if has_tp else None, scale_factor=scale)
            if getattr(self, "device_contours", False) and inst.is_cuda:
                pp.trace_contours(inst.contiguous(), table)     # vertices in tile coordinates, before any sliver crop (pad is background)
            inst = inst[:, :th, :tw]
            inst_h = None      # the instance maps only travel when an image needs the host contour fallback (vertex buffer overflow / no device contours)
            # only the rows / vertices that were written travel (the buffers are sized for the worst case: H*W/10 rows, H*W/2 vertices)
            rows_used = max(int(table["count"].max().item()), 1) if table["count"].numel() else 1
            table = {k: (v[:, :rows_used] if v.dim() > 1 and k != "cont_xy" else v) for k, v in table.items()}
            if "cont_xy" in table:
                used = min(int(table["cont_total"].max().item()), table["cont_xy"].shape[1]) if table["cont_total"].numel() else 0
                table = {**table, "cont_xy": table["cont_xy"][:, : max(used, 1)]}
            tab = {k: v.cpu().numpy() for k, v in table.items()}
            layers = stacks[3][:, :th, :tw, 0].cpu().numpy() if two_task else None
            for j, i in enumerate(group):
                row = {k: v[j] for k, v in tab.items()}
                if "cont_xy" not in row or int(row["cont_total"]) > len(row["cont_xy"]):
                    inst_h = inst.cpu().numpy() if inst_h is None else inst_h
                cols = pp.instance_columns(None if inst_h is None else inst_h[j], row)     # columnar: no Python work per nucleus
                extra = None
                if two_task:   # the layer task of this tile: whole-tile morphology + region contours on the host (OpenCV, like the reference)
                    layer = self.model._proc_ls(layers[j])  # noqa: SLF001
                    extra = (self.model._get_layer_info(layer), layer if keep_maps else None)  # noqa: SLF001
                results.append((i, cols, inst[j] if keep_maps else None, extra))
            group.clear()

        shape = None
        for i in plan.my_tasks:
            x0, y0, x1, y1 = clipped(i)
            s = (y1 - y0, x1 - x0)
            if s[0] <= 0 or s[1] <= 0:
                results.append((i, wm.columns_from_records([]), None, ({}, None) if two_task else None))
                continue
            if s != shape or len(group) >= self.tile_batch:
                flush()
                shape = s
            group.append(i)
        flush()
        results.sort(key=lambda r: r[0])
        return results

    def _wsi_gather(self, results, plan: WsiTilePlan, *, keep_maps: bool):
        """Exchange the per-tile results: object all-gather of the instance records (and of the two-task model's layer results), ONE tensor
        all-gather of the flattened instance maps.  Returns {task index: (inst_dict, inst_map numpy | None, extra)} on every rank."""
        dist, rank, world = self._dist()
        merged = {}
        if dist is None:
            for i, d, m, extra in results:
                merged[i] = (d, m, extra)
            return merged
        gathered = [None] * world
        dist.all_gather_object(gathered, [(i, d, None if m is None else tuple(m.shape), extra) for i, d, m, extra in results])
        maps_by_rank = [None] * world
        if keep_maps:
            flat = [m.reshape(-1) for _, _, m, _ in results if m is not None]
            mine = torch.cat(flat) if flat else torch.empty(0, dtype=torch.int32, device=self.device)
            n = torch.tensor([mine.numel()], dtype=torch.long, device=self.device)
            sizes = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(sizes, n)
            cap = max(int(s.item()) for s in sizes)
            padded = torch.zeros(max(cap, 1), dtype=torch.int32, device=self.device)
            padded[: mine.numel()] = mine
            out = torch.empty(world * max(cap, 1), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(out, padded)
            out = out.view(world, -1)       # stays on the device: the WSI instance map is pasted there (`_wsi_replay`)
            maps_by_rank = [out[r, : int(sizes[r].item())] for r in range(world)]
        for r, part in enumerate(gathered):
            off = 0
            for i, d, shp, extra in part:
                m = None
                if shp is not None and keep_maps:
                    cnt = int(shp[0]) * int(shp[1])
                    m = maps_by_rank[r][off: off + cnt].reshape(shp)
                    off += cnt
                merged[i] = (d, m, extra)
        return merged

    def _wsi_replay(self, merged: dict, plan: WsiTilePlan, *, return_predictions: bool) -> dict:
        """`_process_tile_mode`'s merge loop (:1230-1279) over all tiles in task order -> the output entries of the task (one task), or
        {task name: entries} for the two-task model: its layer regions go through the same margin logic keyed on their boxes, and its
        semantic tile maps are pasted in order (later tiles overwrite, `_update_tile_based_predictions_array` :3160-3212)."""
        store = wm.WsiColumnStore()
        two_task = any(v[2] is not None for v in merged.values())
        layer_store = wm.WsiInstanceStore() if two_task else None
        # the WSI instance map is pasted on the device the tile maps live on (torch ops, no sync per tile) and copied back once
        wsi_pred = torch.zeros((plan.H, plan.W), dtype=torch.int32, device=self.device) if return_predictions else None
        wsi_layer = np.zeros((plan.H, plan.W), dtype=np.uint8) if (return_predictions and two_task) else None
        max_inst = torch.zeros((), dtype=torch.int32, device=self.device)
        lut = None
        for i, (bounds, flag, mode) in enumerate(plan.tasks):
            inst_dict, inst_map, extra = merged.get(i, (wm.columns_from_records([]), None, None))
            if wsi_pred is not None and inst_map is not None:
                inst_map = torch.as_tensor(inst_map).to(self.device)
                need = inst_map.numel() + 2 + wm.PASTE_SCRATCH      # instance ids are bounded by the tile's pixel count
                if lut is None or lut.numel() < need:
                    lut = torch.zeros(need, dtype=torch.int32, device=self.device)
                max_inst = wm.paste_instance_tile_tensor(inst_map, wsi_pred, bounds, max_inst, lut)
            if extra is not None and wsi_layer is not None and extra[1] is not None:
                wm.update_tile_based_predictions_array(extra[1], wsi_layer, bounds, (0, 0), seg_type="semantic")
            if plan.tile_mode:
                store.merge_tile(inst_dict, tile_bounds=bounds, tile_flag=flag, tile_mode=mode, margin=plan.margin, consume=True)
                if extra is not None:
                    layer_store.merge_tile(extra[0], tile_bounds=bounds, tile_flag=flag, tile_mode=mode, margin=plan.margin)
            else:   # `_process_full_wsi` :999-1076: one postproc call over the whole map, nothing to merge
                store.append_all(inst_dict)
                for rec in (extra[0].values() if extra is not None else ()):
                    layer_store._append(rec)  # noqa: SLF001
        out = dict(store.to_info_dict())
        if wsi_pred is not None:
            out["predictions"] = wsi_pred.cpu().numpy()
        if not two_task:
            return out
        layer_out = dict(layer_store.to_info_dict())
        if wsi_layer is not None:
            layer_out["predictions"] = wsi_layer
        tasks = self.model.tasks
        return {tasks[0]: out, tasks[1]: layer_out}

    # ---- entry ---------------------------------------------------------------------------------------------------------------
    def run_wsi(self, image: np.ndarray, *, ioconfig=None, return_predictions: bool = False, return_probabilities: bool = False) -> dict:
        """One in-memory slide / region (H, W, 3) uint8 -> {"box", "centroid", "contours", "prob", "type": object arrays over the merged
        instances in WSI coordinates (absent when nothing was found, like the reference), "coordinates": (N, 4) patch input boxes,
        ["predictions": (H, W) int32 instance map], ["probabilities": [np, hv, tp] WSI head maps (single-rank runs only)]}.  A two-task
        model (HoVerNetPlus) nests the entries under its task names and adds the layer task's "box" / "contours" / "type"
        [/ "predictions": (H, W) uint8 layer map]."""
        image = np.ascontiguousarray(image)
        if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:  # noqa: PLR2004
            msg = "WSI mode expects an in-memory (H, W, 3) uint8 array (slide files are read by WSIReader, outside the B200 path)"
            raise ValueError(msg)
        io = ioconfig or self.ioconfig
        if io is None or io.patch_output_shape is None:
            msg = "an ioconfig with patch_input_shape / patch_output_shape / stride_shape is required for WSI mode"
            raise ValueError(msg)
        dist, rank, world = self._dist()
        plan = WsiTilePlan((image.shape[1], image.shape[0]), io, rank, world)
        self.model._plan_batch = self.batch_size  # noqa: SLF001
        self._stain_failed = None
        import time

        def _sync():
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            return time.perf_counter()

        canvases = []
        t0 = _sync()
        if len(plan.my_patches):
            band = self._wsi_upload(image, plan)
            canvases = self._wsi_infer(band, plan)
            del band
        self._report_stain_failures()
        t1 = _sync()
        results = self._wsi_postproc_tiles(canvases, plan, keep_maps=return_predictions)
        t2 = _sync()
        merged = self._wsi_gather(results, plan, keep_maps=return_predictions)
        t3 = _sync()
        out = self._wsi_replay(merged, plan, return_predictions=return_predictions)
        #: wall-clock split of the last WSI run on this rank (seconds): upload + inference + stitching | tile post-processing |
        #: result exchange | ordered merge replay
        self.wsi_timings = {"infer_s": t1 - t0, "tile_postproc_s": t2 - t1, "gather_s": t3 - t2, "merge_replay_s": time.perf_counter() - t3,
                            "patches_inferred": int(len(plan.my_patches)), "tiles_owned": int(len(plan.my_tasks))}
        out["coordinates"] = plan.in_boxes
        if return_probabilities:
            if world > 1:
                msg = "return_probabilities in WSI mode is only available on a single rank (the head maps are not gathered)"
                raise NotImplementedError(msg)
            out["probabilities"] = [c.cpu().numpy() for c in canvases]
        return out
