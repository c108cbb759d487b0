"""HoVerNet+ (Shephard et al. 2021) for the B200 — mirror of tiatoolbox/models/architecture/hovernetplus.py.

HoVerNetPlus is `HoVerNet(mode="fast")` with a fourth decoder branch `ls` (layer segmentation, hovernetplus.py:107-131), whose head is
argmax(softmax) like `tp` (:455-458), and a two-task post-processing: nuclei with `scale_factor=0.5` (:305 — an 11-tap Sobel and a
3-pixel marker threshold, run on the device by `tiab_hovernet_postproc_scaled`) and the layer map clean-up `_proc_ls` (:135-194).
On the device the fourth branch is one more group of the grouped decoder launches (models/hovernet_plan.py): same launch count as HoVer-Net.

`_proc_ls` / `_get_layer_info` stay on the host: they are whole-image 20x20 morphology + contour tracing on a uint8 class map (OpenCV calls
in the reference), not part of the per-patch arithmetic hot path; the scikit-image call (`morphology.remove_small_objects` on a boolean
mask, a pinned dependency that this image lacks) is restated with `scipy.ndimage.label`."""
from __future__ import annotations

import numpy as np
import torch

from tiatoolbox_b200.models.architecture._params import seeded_state_dict
from tiatoolbox_b200.models.architecture.hovernet import HoVerNet, hovernet_spec, pack_instance_result
from tiatoolbox_b200.models.models_abc import ModelABC


def _remove_small_objects_bool(mask: np.ndarray, max_size: int) -> np.ndarray:
    from scipy import ndimage

    lab = ndimage.label(mask)[0]
    sizes = np.bincount(lab.ravel())
    small = sizes <= max_size
    small[0] = False
    out = mask.copy()
    out[small[lab]] = False
    return out


class HoVerNetPlus(HoVerNet):
    """Same constructor, state-dict keys, `infer_batch` and `postproc` contract as the reference class (hovernetplus.py:23-133)."""

    def __init__(self, num_input_channels: int = 3, num_types: int | None = None, num_layers: int | None = None,
                 nuc_type_dict: dict | None = None, layer_type_dict: dict | None = None) -> None:
        super().__init__(num_input_channels=num_input_channels, num_types=num_types, mode="fast", nuc_type_dict=nuc_type_dict)
        if num_types is None or num_layers is None:
            msg = "HoVerNetPlus needs both num_types and num_layers (its decoder has the four branches tp, np, hv, ls)"
            raise ValueError(msg)
        self.num_layers = num_layers
        self.layer_type_dict = layer_type_dict
        self.tasks = ["nuclei_segmentation", "layer_segmentation"]
        self.class_dict = {self.tasks[0]: nuc_type_dict, self.tasks[1]: layer_type_dict}
        self._sd = seeded_state_dict(hovernet_spec(num_input_channels, num_types, "fast", num_layers), seed=0)


    def _build_plan(self, batch, device, shape):
        from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan

        size = None if shape is None else int(shape[0])
        return HoVerNetPlan(self._sd, batch=batch, device=device, mode="fast", num_types=self.num_types, in_size=size,
                            precision=self.precision, num_layers=self.num_layers)

    @staticmethod
    def infer_batch(model: "HoVerNetPlus", batch_data, *, device) -> tuple[np.ndarray, ...]:
        """hovernetplus.py:413-461: (np, hv, tp, ls) float32 numpy arrays, NHWC."""
        return tuple(o.cpu().numpy() for o in HoVerNet.infer_batch_device(model, batch_data, device=device))

    # ---- post-processing ------------------------------------------------------------------------------------------------
    @staticmethod
    def _proc_ls(ls_map: np.ndarray) -> np.ndarray:
        """Layer map clean-up (hovernetplus.py:135-194): drop epithelium islands < 20 000 px, close + open every class with a 20x20 box."""
        import cv2

        ls_map = np.around(np.squeeze(ls_map)).astype("uint8")
        min_size, kernel_size = 20000, 20
        epith_all = np.where(ls_map >= 2, 1, 0).astype("uint8") > 0  # noqa: PLR2004
        mask = np.where(ls_map >= 1, 1, 0).astype("uint8")
        epith_mask = _remove_small_objects_bool(epith_all, max_size=min_size - 1).astype("uint8")
        epith_edited = (epith_mask * ls_map).astype("uint8")
        epith_edited_open = np.zeros_like(epith_edited).astype("uint8")
        box = np.ones((kernel_size, kernel_size))
        for i in [3, 2, 4]:
            tmp = np.where(epith_edited == i, 1, 0).astype("uint8")
            ep_open = cv2.morphologyEx(cv2.morphologyEx(tmp, cv2.MORPH_CLOSE, box), cv2.MORPH_OPEN, box)
            epith_edited_open[ep_open == 1] = i
        mask_open = cv2.morphologyEx(cv2.morphologyEx(mask, cv2.MORPH_CLOSE, box), cv2.MORPH_OPEN, box).astype("uint8")
        out = mask_open.copy()
        for i in range(2, 5):
            out[epith_edited_open == i] = i
        return out.astype("uint8")

    @staticmethod
    def _get_layer_info(pred_layer: np.ndarray, offset: tuple[int, int] = (0, 0)) -> dict:
        """Layer regions as contours (hovernetplus.py:197-264): {count: {"box", "contours", "type"}}."""
        import cv2

        offset = np.asarray(offset)
        info, count = {}, 1
        for type_class in [v for v in np.unique(pred_layer) if v != 0]:
            layer = np.where(pred_layer == type_class, 1, 0).astype("uint8")
            rows, cols = np.any(layer, axis=1), np.any(layer, axis=0)
            r0, r1 = np.where(rows)[0][[0, -1]]
            c0, c1 = np.where(cols)[0][[0, -1]]
            bounding_box = np.array([c0, r0, c1 + 1, r1 + 1])          # utils/misc.py get_bounding_box
            contours, _ = cv2.findContours(layer, cv2.RETR_TREE, cv2.CHAIN_APPROX_NONE)
            for cont in contours:
                sq = np.squeeze(cont)
                if sq.shape[0] < 3 or len(sq.shape) != 2:  # noqa: PLR2004
                    continue
                box = bounding_box.copy()
                box[:2] = box[:2] + offset
                box[2:] = box[2:] + offset
                info[count] = {"box": box, "contours": cont[:, 0, :] + offset, "type": type_class}
                count += 1
        return info

    def postproc(self, raw_maps, offset: tuple[int, int] = (0, 0)) -> tuple[dict, dict]:
        """hovernetplus.py:267-410 for one image: (nuclei task dict, layer task dict)."""
        from tiatoolbox_b200.models import postproc as pp

        np_map, hv_map, tp_map, ls_map = (np.asarray(a) for a in raw_maps)
        dev = self._device
        t = [torch.as_tensor(a, dtype=torch.float32).to(dev)[None] for a in (np_map, hv_map, tp_map)]
        inst, table = pp.proc_np_hv_batch(t[0], t[1], t[2], scale_factor=0.5)
        pred_inst = inst[0].cpu().numpy()
        info = pp.instance_info(pred_inst, {k: v[0].cpu().numpy() for k, v in table.items()}, offset)
        nuclei = pack_instance_result(self.tasks[0], pred_inst, info)
        pred_layer = self._proc_ls(ls_map)
        layer_info = self._get_layer_info(pred_layer, offset)
        if not layer_info:
            empty = np.empty(0)
            cols = {"box": empty, "contours": empty, "type": empty}
        else:
            cols = {}
            for key in ("box", "contours", "type"):
                col = np.empty(len(layer_info), dtype=object)
                for i, rec in enumerate(layer_info.values()):
                    col[i] = rec[key]
                cols[key] = col
        layer = {"task_type": self.tasks[1], "predictions": pred_layer, "info_dict": cols, "seg_type": "semantic"}
        return nuclei, layer


__all__ = ["HoVerNetPlus", "ModelABC"]
