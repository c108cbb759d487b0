"""Patch-level inference engines on the B200 (API mirror of tiatoolbox/models/engine)."""
from tiatoolbox_b200.models.engine.engine_abc import EngineABC
from tiatoolbox_b200.models.engine.nucleus_instance_segmentor import MultiTaskSegmentor, NucleusInstanceSegmentor
from tiatoolbox_b200.models.engine.patch_predictor import PatchPredictor
from tiatoolbox_b200.models.engine.semantic_segmentor import SemanticSegmentor

__all__ = ["EngineABC", "MultiTaskSegmentor", "NucleusInstanceSegmentor", "PatchPredictor", "SemanticSegmentor"]
