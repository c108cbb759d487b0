"""PatchPredictor (mirror of tiatoolbox/models/engine/patch_predictor.py): `resnet18-kather100k` & co."""
from __future__ import annotations

import numpy as np
import torch

from tiatoolbox_b200.models.engine.engine_abc import EngineABC, cast_to_min_dtype


class PatchPredictor(EngineABC):
    """run(...) -> {"predictions": (N,), ["probabilities": (N, C)], ["labels"]} (patch_predictor.py:377-380, 448-537)."""

    def infer_patches(self, patches: np.ndarray) -> torch.Tensor:
        outs = [type(self.model).infer_batch_device(self.model, x, device=self.device).clone() for x in self._batches(patches)]
        c = getattr(self.model, "num_classes", 1)
        return torch.cat(outs) if outs else torch.empty((0, c), device=self.device)

    def run(self, images, *, patch_mode: bool = True, patch_input_shape=None, stride_shape=None, ioconfig=None, **kwargs):
        """Patch mode: EngineABC.run.  `patch_mode=False` with an in-memory (H, W, 3) uint8 slide / region: the reference's WSI flow
        (engine_abc.py `infer_wsi`: patches on the `PatchExtractor.get_coordinates` grid, 255-padded outside the image, dataset_abc.py:419-431)
        -> the patch-mode results plus `"coordinates"` (N, 4) [start_x, start_y, end_x, end_y].  Slide files are WSIReader I/O (out of scope)."""
        if patch_mode or not (isinstance(images, np.ndarray) and images.ndim == 3):  # noqa: PLR2004
            return super().run(images, patch_mode=patch_mode, patch_input_shape=patch_input_shape, ioconfig=ioconfig, **kwargs)
        from tiatoolbox_b200.tools.patchextraction import PatchExtractor

        io = ioconfig or self.ioconfig
        shape = patch_input_shape or (io.patch_input_shape if io is not None else None)
        if shape is None:
            msg = "patch_input_shape (or an ioconfig) is required to tile a slide"
            raise ValueError(msg)
        stride = stride_shape or (io.stride_shape if io is not None and patch_input_shape is None else shape)
        image = np.ascontiguousarray(images)
        if image.dtype != np.uint8 or image.shape[2] != 3:  # noqa: PLR2004
            msg = "WSI mode expects an in-memory (H, W, 3) uint8 array"
            raise ValueError(msg)
        h, w = image.shape[:2]
        coords = PatchExtractor.get_coordinates(image_shape=(w, h), patch_input_shape=tuple(int(v) for v in shape),
                                                stride_shape=tuple(int(v) for v in stride))
        pw, ph = int(shape[0]), int(shape[1])
        patches = np.full((len(coords), ph, pw, 3), 255, np.uint8)
        for k, (x0, y0, x1, y1) in enumerate(coords):
            xs, ys, xe, ye = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
            if xe > xs and ye > ys:
                patches[k, ys - y0: ye - y0, xs - x0: xe - x0] = image[ys:ye, xs:xe]
        out = super().run(patches, patch_mode=True, **{k: v for k, v in kwargs.items() if k not in ("masks", "input_resolutions")})
        out["coordinates"] = coords
        return out

    predict = run

    def post_process_patches(self, raw: torch.Tensor, n_total: int) -> dict:
        probs = self._all_gather_rows(raw, n_total).cpu().numpy()
        preds = cast_to_min_dtype(np.asarray(self.model.postproc_func(probs)))
        out = {"predictions": preds}
        if self.return_probabilities:
            out["probabilities"] = probs
        return out
