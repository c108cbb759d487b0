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

    def post_process_patches(self, raw: torch.Tensor, n_total: int) -> dict:
        probs = self._all_gather_rows(raw, n_total).cpu().numpy()
        preds = cast_to_min_dtype(np.asarray(self.model.postproc_func(probs)))
        out = {"predictions": preds}
        if self.return_probabilities:
            out["probabilities"] = probs
        return out
