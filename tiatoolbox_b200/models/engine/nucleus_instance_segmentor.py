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

    #: rows of the instance table / contour vertices copied to the host per image in the common case (a 164-px HoVer-Net map holds a few
    #: dozen nuclei); an image that needs more is fetched in full when its batch is finished
    table_rows_fast = 512
    contour_vertices_fast = 4096

    def infer_patches(self, patches: np.ndarray):
        """GPU: stain -> network -> post-processing (+ contour tracing) per batch, all asynchronous.  Host: the instance dictionaries of
        batch i are assembled while the GPU already works on batch i + 1 (results travel through pinned buffers; with
        `device_contours=False` the host also traces the contours with cv2 there)."""
        insts, infos, heads, layers, inst_host = [], [], [], [], []
        pending = None
        two_task = getattr(self.model, "num_layers", None) is not None      # HoVerNetPlus: nuclei (scale 0.5) + layer segmentation
        scale = 0.5 if two_task else 1.0
        self.d2h_bytes = 0

        def finish(p):
            ev, inst_h, table_h, table_d = p
            ev.synchronize()
            inst_np = inst_h.numpy().copy()           # the pinned buffer is reused two batches later
            inst_host.append(inst_np)
            t = {k: v.numpy() for k, v in table_h.items()}
            need_rows = int(t["count"].max()) if len(t["count"]) else 0
            need_xy = int(t["cont_total"].max()) if "cont_total" in t and len(t["cont_total"]) else 0
            if need_rows > t["id"].shape[1] or ("cont_xy" in t and need_xy > t["cont_xy"].shape[1]):
                t = {k: v.cpu().numpy() for k, v in table_d.items()}      # rare: an unusually crowded image, fetch the whole table
                self.d2h_bytes += sum(v.numel() * v.element_size() for v in table_d.values())
            for i in range(len(t["count"])):
                infos.append(pp.instance_info(inst_np[i], {key: v[i] for key, v in t.items()}, with_contours=self.with_contours))

        for b, x in enumerate(self._batches(patches)):
            outs = type(self.model).infer_batch_device(self.model, x, device=self.device)
            np_map, hv_map = outs[0], outs[1]
            tp_map = outs[2] if len(outs) > 2 else None
            inst, table = pp.proc_np_hv_batch(np_map, hv_map, tp_map, scale_factor=scale)
            if self.device_contours and self.with_contours:
                pp.trace_contours(inst, table)
            insts.append(inst)
            if two_task:   # the layer head is a class-index map; its clean-up is whole-map morphology on the host (hovernetplus.py:135-194)
                layers.append(outs[3][..., 0].to(torch.uint8).cpu().numpy())
            slot = b & 1
            inst_h = self._to_host_async(("inst", slot), inst)
            table_h = {}
            for k, v in table.items():
                cap = self.contour_vertices_fast if k == "cont_xy" else (self.table_rows_fast if v.dim() > 1 else None)
                table_h[k] = self._to_host_async(("tab", k, slot), v if cap is None or v.shape[1] <= cap else v[:, :cap])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            if self.return_probabilities:
                heads.append([o.clone() for o in outs])
            if pending is not None:
                finish(pending)
            pending = (ev, inst_h, table_h, table)
        if pending is not None:
            finish(pending)
        self._inst_host = inst_host
        if two_task:
            return insts, infos, heads, [type(self.model)._proc_ls(m) for batch in layers for m in batch]  # noqa: SLF001
        return insts, infos, heads

    def post_process_patches(self, raw, n_total: int) -> dict:
        insts, infos, heads = raw[:3]
        dist, rank, world = self._dist()
        if dist is None and insts:   # single rank: the instance maps already came back through the pinned buffers
            inst_all = np.concatenate(self._inst_host)
        else:
            local = torch.cat(insts) if insts else torch.empty((0, 1, 1), dtype=torch.int32, device=self.device)
            inst_all = self._all_gather_rows(local, n_total).cpu().numpy()
            self.d2h_bytes += inst_all.nbytes
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, infos)
            infos = [x for part in gathered for x in part]
        out = {"predictions": inst_all}
        from tiatoolbox_b200.models.architecture.hovernet import pack_instance_result

        cols = {key: [] for key in ("box", "centroid", "contours", "prob", "type")}
        for i, info in enumerate(infos):
            packed = pack_instance_result("nuclei_segmentation", inst_all[i], info)["info_dict"]
            for key in cols:
                cols[key].append(packed[key])
        out.update(cols)
        probabilities = None
        if self.return_probabilities and heads:
            probabilities = [self._all_gather_rows(torch.cat([h[j] for h in heads]), n_total).cpu().numpy() for j in range(len(heads[0]))]
        if len(raw) > 3:   # two tasks: results nested under the task names (multi_task_segmentor.py:1631-1661)
            layer_maps = raw[3]
            local = torch.from_numpy(np.stack(layer_maps)) if layer_maps else torch.empty((0, 1, 1), dtype=torch.uint8)
            layer_all = self._all_gather_rows(local.to(self.device), n_total).cpu().numpy()
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


class NucleusInstanceSegmentor(MultiTaskSegmentor):
    """Thin subclass kept for API compatibility (nucleus_instance_segmentor.py:103-174)."""
