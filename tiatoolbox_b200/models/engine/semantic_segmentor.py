"""SemanticSegmentor (mirror of tiatoolbox/models/engine/semantic_segmentor.py, patch mode): `fcn-tissue_mask`."""
from __future__ import annotations

import numpy as np
import torch

from tiatoolbox_b200.models.engine.engine_abc import EngineABC, cast_to_min_dtype


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
