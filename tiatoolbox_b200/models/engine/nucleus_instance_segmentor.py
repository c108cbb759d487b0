"""NucleusInstanceSegmentor / MultiTaskSegmentor (mirror of tiatoolbox/models/engine/multi_task_segmentor.py,
nucleus_instance_segmentor.py; patch mode): `hovernet_fast-pannuke`.

The reference runs HoVerNet.postproc per patch in a Python loop on the CPU (multi_task_segmentor.py:825-828);
here the watershed + instance table run on the device for the whole batch, the instance maps are all-gathered
over NCCL, and only contour tracing (cv2.findContours) remains on the host."""
from __future__ import annotations

import numpy as np
import torch

from tiatoolbox_b200.models import postproc as pp
from tiatoolbox_b200.models.engine.engine_abc import EngineABC
from tiatoolbox_b200.models.engine.wsi import WsiTileModeMixin


class MultiTaskSegmentor(WsiTileModeMixin, EngineABC):
    """run(...) -> single task flattened (multi_task_segmentor.py:1697-1704):
    {"predictions": (N, h, w) int32, "box", "centroid", "contours", "prob", "type": length-N lists,
     ["probabilities": [np, hv, tp]]}."""

    with_contours = True
    #: trace the instance contours on the device (csrc/contour_math.cuh: equal to cv2.findContours on fuzzed masks on the CPU and on
    #: the B200, profiles/r1h_gpu_tests_contours.log); False = cv2.findContours per nucleus on the host (the r1g behaviour)
    device_contours = True

    def run(self, images, *, patch_mode: bool = True, ioconfig=None, return_predictions=None, **kwargs):
        """Patch mode: EngineABC.run.  `patch_mode=False` with in-memory slides — one (H, W, 3) uint8 array or a list of them — runs the
        WSI data flow of multi_task_segmentor.py:477-731, 836-1287 on the device (`WsiTileModeMixin.run_wsi`) and returns one result
        dict per slide (a single array gives a single dict).  `return_predictions` follows the reference: a tuple of per-task flags
        (one task here) or a bool."""
        if patch_mode:
            return super().run(images, patch_mode=True, ioconfig=ioconfig, **kwargs)
        if kwargs.get("output_type", "dict").lower() != "dict":
            msg = "only output_type='dict' is produced on the B200 path (zarr / AnnotationStore are result persistence)"
            raise NotImplementedError(msg)
        single = isinstance(images, np.ndarray) and images.ndim == 3  # noqa: PLR2004
        slides = [images] if single else list(images)
        if not slides or not all(isinstance(s, np.ndarray) for s in slides):
            msg = "WSI mode takes in-memory (H, W, 3) uint8 arrays (slide files are read by WSIReader, outside the B200 path)"
            raise NotImplementedError(msg)
        for k, v in kwargs.items():
            if k not in ("masks", "input_resolutions", "patch_input_shape", "save_dir", "overwrite", "output_type"):
                setattr(self, k, v)
        want_pred = bool(return_predictions[0]) if isinstance(return_predictions, (tuple, list)) else bool(return_predictions)
        outs = [self.run_wsi(s, ioconfig=ioconfig, return_predictions=want_pred, return_probabilities=bool(self.return_probabilities))
                for s in slides]
        return outs[0] if single else outs

    predict = run

    #: rows of the instance table / contour vertices that travel per image in the common case (a 164-px HoVer-Net map holds a few
    #: dozen nuclei); an image that needs more is fetched in full from its owner's device when its batch is finished
    table_rows_fast = 512
    contour_vertices_fast = 4096

    _ICOLS = (("id", 1), ("box", 4), ("type", 1), ("cont_off", 1), ("cont_len", 1))     # int32 words per table row
    _FCOLS = (("centroid", 2), ("prob", 1))                                              # float64 values per table row

    def _record_words(self, hw, rows: int, verts: int) -> int:
        n = 2 + 2 * rows * sum(w for _, w in self._FCOLS) + hw[0] * hw[1] + rows * sum(w for _, w in self._ICOLS) + 2 * verts
        return n + (n & 1)            # even number of int32 words: every record (and its float64 block) stays 8-byte aligned

    def _pack_batch(self, inst: torch.Tensor, table: dict, rows: int, verts: int) -> torch.Tensor:
        """One int32 record per image: [count, cont_total | float64 columns (viewed as int32 pairs) capped at `rows` | instance map |
        int32 table columns capped at `rows` | contour vertices capped at `verts`], padded to `batch_size` records (count = -1 marks
        padding).  This is what crosses NVLink (one all_gather_into_tensor per batch) and PCIe (one pinned D2H per batch)."""
        n = inst.shape[0]
        dev = inst.device
        has_c = "cont_xy" in table

        def capped(t, cap):
            t = t[:, :cap]
            if t.shape[1] < cap:
                t = torch.cat([t, torch.zeros((n, cap - t.shape[1], *t.shape[2:]), dtype=t.dtype, device=dev)], 1)
            return t.reshape(n, -1)

        zeros = torch.zeros((n,), dtype=torch.int32, device=dev)
        parts = [table["count"].to(torch.int32)[:, None], (table["cont_total"] if has_c else zeros)[:, None]]
        for k, _ in self._FCOLS:
            parts.append(capped(table[k], rows).contiguous().view(torch.int32))
        parts.append(inst.reshape(n, -1))
        for k, w in self._ICOLS:
            parts.append(capped(table[k], rows) if (has_c or not k.startswith("cont_")) else torch.zeros((n, rows * w), dtype=torch.int32, device=dev))
        parts.append(capped(table["cont_xy"], verts) if has_c else torch.zeros((n, 2 * verts), dtype=torch.int32, device=dev))
        rec = torch.cat(parts, 1)
        width = self._record_words(tuple(inst.shape[1:]), rows, verts)
        if rec.shape[1] < width:
            rec = torch.cat([rec, torch.zeros((n, width - rec.shape[1]), dtype=torch.int32, device=dev)], 1)
        if n < self.batch_size:
            pad = torch.zeros((self.batch_size - n, width), dtype=torch.int32, device=dev)
            pad[:, 0] = -1
            rec = torch.cat([rec, pad])
        return rec.contiguous()

    def _unpack(self, rec: np.ndarray, hw: tuple[int, int], rows: int, verts: int) -> tuple[np.ndarray, dict]:
        """Views into gathered records (R, L): instance maps (R, H, W) and the table columns keyed like `proc_np_hv_batch`'s table."""
        h, w = hw
        t = {"count": rec[:, 0], "cont_total": rec[:, 1]}
        o = 2
        for k, wd in self._FCOLS:
            v = rec[:, o:o + 2 * rows * wd].view(np.float64)
            t[k] = v.reshape(-1, rows, wd) if wd > 1 else v
            o += 2 * rows * wd
        inst = rec[:, o:o + h * w].reshape(-1, h, w)
        o += h * w
        for k, wd in self._ICOLS:
            v = rec[:, o:o + rows * wd]
            t[k] = v.reshape(-1, rows, wd) if wd > 1 else v
            o += rows * wd
        t["cont_xy"] = rec[:, o:o + 2 * verts].reshape(-1, verts, 2)
        return inst, t

    def infer_patches(self, patches: np.ndarray):
        """GPU: stain -> network -> post-processing (+ contour tracing) per batch, all asynchronous; with several ranks every batch ends in
        ONE all-gather (NCCL) of the fixed-size per-image records (instance map + capped instance table + contour vertices), so that the
        exchange and the device -> host copy of batch i overlap the compute of batch i + 1.  Host: while the GPU works on batch i + 1 the
        instance dictionaries of THIS rank's images of batch i are assembled; the records of the other ranks' images are only stored
        (compact copies) and turned into Python objects when they are accessed (`_ShardedColumn`).  Per-rank host work therefore does not
        grow with the world size."""
        dist, rank, world = self._dist()
        n_local = len(patches)
        n_total = int(getattr(self, "_n_total", None) or n_local) if dist is not None else n_local
        per = -(-n_total // world) if dist is not None else n_local
        rounds = -(-per // self.batch_size) if per else 0
        two_task = getattr(self.model, "num_layers", None) is not None      # HoVerNetPlus: nuclei (scale 0.5) + layer segmentation
        scale = 0.5 if two_task else 1.0
        self.d2h_bytes = 0
        rows, verts = int(self.table_rows_fast), int(self.contour_vertices_fast)
        heads, layers, insts_dev = [], [], []
        st = {"inst_all": None, "infos": [None] * n_local, "remote": {}, "overflow": []}
        pending = None
        use_contours = self.device_contours and self.with_contours

        import time as _time

        tm = {"wait_s": 0.0, "finish_s": 0.0, "issue_s": 0.0, "rounds": rounds}     # host-side split of this call (diagnostics: `last_timings`)
        self.last_timings = tm

        def finish(p):
            ev, host, table_d, base, hw = p
            _t0 = _time.perf_counter()
            ev.synchronize()
            _t1 = _time.perf_counter()
            tm["wait_s"] += _t1 - _t0
            try:
                _finish_body(host, table_d, base, hw)
            finally:
                tm["finish_s"] += _time.perf_counter() - _t1

        def _finish_body(host, table_d, base, hw):
            rec = host.numpy()
            inst_v, t = self._unpack(rec, hw, rows, verts)
            if st["inst_all"] is None:     # every image of the job is written below, so the recycled block needs no clearing
                from tiatoolbox_b200.models.engine.engine_abc import _RESULT_POOL

                st["inst_all"] = _RESULT_POOL.take((n_total, *hw), np.int32)
            for r in range(world if dist is not None else 1):
                blk = slice(r * self.batch_size, (r + 1) * self.batch_size)
                cnt = t["count"][blk]
                nv = int((cnt >= 0).sum())
                if nv == 0:
                    continue
                g0 = r * per + base
                st["inst_all"][g0:g0 + nv] = inst_v[blk][:nv]
                if r == rank or dist is None:
                    tr = {k: v[blk][:nv] for k, v in t.items()}
                    need_rows = int(cnt[:nv].max())
                    need_xy = int(tr["cont_total"].max()) if use_contours else 0
                    if need_rows > rows or need_xy > verts:       # rare: an unusually crowded image, fetch the whole table of the batch
                        tr = {k: v.cpu().numpy()[:nv] for k, v in table_d.items()}
                        self.d2h_bytes += sum(v.numel() * v.element_size() for v in table_d.values())
                        st["overflow"].extend(range(g0, g0 + nv))
                    elif not use_contours:
                        tr = {k: v for k, v in tr.items() if not k.startswith("cont_")}
                    for i in range(nv):
                        st["infos"][base + i] = pp.instance_info(st["inst_all"][g0 + i], {key: v[i] for key, v in tr.items()},
                                                                 with_contours=self.with_contours)
                else:                                              # another rank's images: keep a compact copy, materialise on access
                    ru = max(int(cnt[:nv].max()), 0)
                    vu = max(int(t["cont_total"][blk][:nv].max()), 0)
                    comp = {"count": cnt[:nv].copy(), "cont_total": t["cont_total"][blk][:nv].copy()}
                    for k, _ in self._ICOLS + self._FCOLS:
                        comp[k] = t[k][blk][:nv, :min(ru, rows)].copy()
                    comp["cont_xy"] = t["cont_xy"][blk][:nv, :min(vu, verts)].copy()
                    if not use_contours:
                        comp = {k: v for k, v in comp.items() if not k.startswith("cont_")}
                    for i in range(nv):
                        st["remote"][g0 + i] = (comp, i, ru > rows or vu > verts)

        batches = self._batches(patches)
        for b in range(rounds):
            _ti = _time.perf_counter()
            x = next(batches, None)
            if x is not None:
                outs = type(self.model).infer_batch_device(self.model, x, device=self.device)
                np_map, hv_map = outs[0], outs[1]
                tp_map = outs[2] if len(outs) > 2 else None
                inst, table = pp.proc_np_hv_batch(np_map, hv_map, tp_map, scale_factor=scale)
                if use_contours:
                    pp.trace_contours(inst, table)
                hw = tuple(inst.shape[1:])
                st["hw"] = hw
                if two_task:   # the layer head is a class-index map; its clean-up is whole-map morphology on the host (hovernetplus.py:135-194)
                    layers.append(outs[3][..., 0].to(torch.uint8).cpu().numpy())
                if self.return_probabilities:
                    heads.append([o.clone() for o in outs])
                rec = self._pack_batch(inst, table, rows, verts)
            else:   # this rank's shard is exhausted (last rank of an uneven split): an all-padding record keeps the ranks in lock step
                hw = st.get("hw") or self._empty_round_shape(dist)
                table = None
                rec = torch.zeros((self.batch_size, self._record_words(hw, rows, verts)), dtype=torch.int32, device=self.device)
                rec[:, 0] = -1
            if dist is not None:
                gathered = torch.empty((world * self.batch_size, rec.shape[1]), dtype=torch.int32, device=rec.device)
                dist.all_gather_into_tensor(gathered, rec)
                rec = gathered
            host, ev = self._to_host_async(("rec", b & 1), rec, side_stream=True)
            tm["issue_s"] += _time.perf_counter() - _ti
            if pending is not None:
                finish(pending)
            pending = (ev, host, table, b * self.batch_size, hw)
        if pending is not None:
            finish(pending)
        if st["inst_all"] is None:
            st["inst_all"] = np.zeros((n_total, 1, 1), np.int32)
        self._state = st
        infos = st["infos"]
        if two_task:
            return insts_dev, infos, heads, [type(self.model)._proc_ls(m) for batch in layers for m in batch]  # noqa: SLF001
        return insts_dev, infos, heads

    def _empty_round_shape(self, dist) -> tuple[int, int]:
        """Output map size for a rank that never ran a batch (more ranks than batches): taken from the model's ioconfig, else agreed on
        through the model's plan geometry (HoVer-Net fast: 164)."""
        io = getattr(self, "ioconfig", None)
        shp = getattr(io, "patch_output_shape", None)
        if shp:
            return int(shp[1]), int(shp[0])
        return 164, 164

    def post_process_patches(self, raw, n_total: int) -> dict:
        infos, heads = raw[1], raw[2]
        dist, rank, world = self._dist()
        st = self._state
        inst_all = st["inst_all"]
        from tiatoolbox_b200.models.architecture.hovernet import pack_instance_result

        task = self.model.tasks[0] if hasattr(self.model, "tasks") else "nuclei_segmentation"
        keys = ("box", "centroid", "contours", "prob", "type")
        lo, hi = self.shard_bounds(n_total, rank, world) if dist is not None else (0, n_total)
        local = {key: [] for key in keys}
        for i, info in enumerate(infos):
            packed = pack_instance_result(task, inst_all[lo + i], info or {})["info_dict"]
            for key in keys:
                local[key].append(packed[key])
        out = {"predictions": inst_all}
        if dist is None:
            out.update(local)
        else:
            # images of other ranks whose table / contour vertices did not fit the fixed-size record: their owners send the packed columns
            flag = torch.tensor([len(st["overflow"])], dtype=torch.int64, device=self.device)
            dist.all_reduce(flag)
            extra = {}
            if int(flag.item()):
                mine = {g: {key: local[key][g - lo] for key in keys} for g in st["overflow"]}
                parts = [None] * world
                dist.all_gather_object(parts, mine)
                for part in parts:
                    extra.update(part)
            store = _RemoteStore(inst_all, st["remote"], extra, task, self.with_contours)
            out.update({key: _ShardedColumn(key, n_total, lo, hi, local[key], store) for key in keys})
        probabilities = None
        if self.return_probabilities:
            n_heads = torch.tensor([len(heads[0]) if heads else 0], dtype=torch.int64, device=self.device)
            if dist is not None:
                dist.all_reduce(n_heads, op=dist.ReduceOp.MAX)
            probabilities = []
            for j in range(int(n_heads.item())):
                loc = torch.cat([h[j] for h in heads]) if heads else torch.empty((0,), dtype=torch.float32, device=self.device)
                probabilities.append(self._all_gather_rows(loc, n_total).cpu().numpy())
        if len(raw) > 3:   # two tasks: results nested under the task names (multi_task_segmentor.py:1631-1661)
            layer_maps = raw[3]
            local_l = torch.from_numpy(np.stack(layer_maps)) if layer_maps else torch.empty((0, 1, 1), dtype=torch.uint8)
            layer_all = self._all_gather_rows(local_l.to(self.device), n_total).cpu().numpy()
            linfo = {key: [] for key in ("box", "contours", "type")}
            for m in layer_all:
                info = type(self.model)._get_layer_info(m)  # noqa: SLF001
                for key in linfo:
                    col = np.empty(len(info), dtype=object)
                    for i, rec in enumerate(info.values()):
                        col[i] = rec[key]
                    linfo[key].append(col if len(info) else np.empty(0))
            tasks = self.model.tasks
            out = {tasks[0]: out, tasks[1]: {"predictions": layer_all, **linfo}}
        if probabilities is not None:
            out["probabilities"] = probabilities
        return out


class _RemoteStore:
    """Per-image instance records of the OTHER ranks' shards, kept as the compact arrays that came out of the all-gather; the reference's
    per-image object arrays (`box`, `centroid`, `contours`, `prob`, `type`) are built from them on first access and cached."""

    def __init__(self, inst_all, remote: dict, extra: dict, task: str, with_contours: bool):
        self.inst_all, self.remote, self.extra, self.task, self.with_contours = inst_all, remote, extra, task, with_contours
        self.cache = {}

    def patch(self, g: int) -> dict:
        if g in self.cache:
            return self.cache[g]
        if g in self.extra:
            cols = self.extra[g]
        else:
            from tiatoolbox_b200.models.architecture.hovernet import pack_instance_result

            comp, i, _ = self.remote[g]
            row = {k: v[i] for k, v in comp.items()}
            info = pp.instance_info(self.inst_all[g], row, with_contours=self.with_contours)
            cols = pack_instance_result(self.task, self.inst_all[g], info)["info_dict"]
        self.cache[g] = cols
        return cols


class _ShardedColumn:
    """A result column (`out["box"]` ...) of a multi-rank run: a sequence with one entry per image of the WHOLE job.  Entries of this
    rank's shard are ready-made objects; entries of other ranks' shards are materialised from the gathered records on access
    (`materialize()` returns a plain list of all of them)."""

    def __init__(self, key, n_total, lo, hi, local, store):
        self.key, self.n, self.lo, self.hi, self.local, self.store = key, n_total, lo, hi, local, store

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        if self.lo <= i < self.hi:
            return self.local[i - self.lo]
        return self.store.patch(i)[self.key]

    def __iter__(self):
        return (self[i] for i in range(self.n))

    def materialize(self) -> list:
        return list(self)


class NucleusInstanceSegmentor(MultiTaskSegmentor):
    """Thin subclass kept for API compatibility (nucleus_instance_segmentor.py:103-174)."""
