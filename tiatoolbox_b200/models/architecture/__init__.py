"""Model registry for the hot-path models (mirrors tiatoolbox/models/architecture/__init__.py:70-178 and the
entries of tiatoolbox/data/pretrained_model.yaml:15-28, 572-592, 642-672, 758-794).

The reference downloads weights from Hugging Face; there is no network here, so `weights=` takes a local
`.pth` (a tiatoolbox checkpoint loads unchanged — the state-dict keys are identical) and, when it is
None, the model keeps its seeded random weights and a warning is logged."""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from pathlib import Path

logger = logging.getLogger(__name__)


@dataclass
class IOConfig:
    """The fields of io_config.IOPatchPredictorConfig / IOSegmentorConfig / IOInstanceSegmentorConfig the
    patch-mode path uses (engine/io_config.py:15-461)."""

    patch_input_shape: tuple[int, int]
    stride_shape: tuple[int, int]
    input_resolutions: list = field(default_factory=list)
    output_resolutions: list = field(default_factory=list)
    patch_output_shape: tuple[int, int] | None = None
    save_resolution: dict | None = None
    margin: int | None = None
    tile_shape: tuple[int, int] | None = None
    ignore_index: int | None = None


PRETRAINED_INFO = {
    "resnet18-kather100k": {
        "class": ("vanilla", "CNNModel"), "kwargs": {"backbone": "resnet18", "num_classes": 9},
        "ioconfig": IOConfig((224, 224), (224, 224), [{"resolution": 0.5, "units": "mpp"}]),
    },
    "fcn-tissue_mask": {
        "class": ("unet", "UNetModel"),
        "kwargs": {"num_input_channels": 3, "num_output_channels": 2, "encoder": "resnet50", "decoder_block": [3]},
        "ioconfig": IOConfig((1024, 1024), (450, 450), [{"units": "mpp", "resolution": 2.0}], [{"units": "mpp", "resolution": 2.0}],
                             (512, 512), {"units": "mpp", "resolution": 8.0}, ignore_index=0),
    },
    "hovernet_fast-pannuke": {
        "class": ("hovernet", "HoVerNet"),
        "kwargs": {"num_types": 6, "mode": "fast",
                   "nuc_type_dict": {0: "Background", 1: "Neoplastic", 2: "Inflammatory", 3: "Connective", 4: "Dead",
                                     5: "Non-Neoplastic Epithelial"}},
        "ioconfig": IOConfig((256, 256), (164, 164), [{"units": "mpp", "resolution": 0.25}], [{"units": "mpp", "resolution": 0.25}] * 3,
                             (164, 164), {"units": "mpp", "resolution": 0.25}, margin=128, tile_shape=(1024, 1024), ignore_index=0),
    },
    "hovernetplus-oed": {
        "class": ("hovernetplus", "HoVerNetPlus"),
        "kwargs": {"num_types": 3, "num_layers": 5, "nuc_type_dict": {0: "Background", 1: "Other", 2: "Epithelial"},
                   "layer_type_dict": {0: "Background", 1: "Other Tissue", 2: "Basal Epithelium", 3: "(Core) Epithelium", 4: "Keratin"}},
        "ioconfig": IOConfig((256, 256), (164, 164), [{"units": "mpp", "resolution": 0.5}], [{"units": "mpp", "resolution": 0.5}] * 4,
                             (164, 164), {"units": "mpp", "resolution": 0.5}, margin=128, tile_shape=(2048, 2048), ignore_index=0),
    },
}


def get_pretrained_model(pretrained_model: str | None = None, pretrained_weights: str | Path | None = None, **kwargs):  # noqa: ARG001
    """Return (model, ioconfig) for a registry name (architecture/__init__.py:70-178)."""
    import importlib

    if not isinstance(pretrained_model, str):
        msg = "pretrained_model must be a string."
        raise ValueError(msg)  # noqa: TRY004
    if pretrained_model not in PRETRAINED_INFO:
        msg = f"Pretrained model `{pretrained_model}` does not exist (B200 path: {sorted(PRETRAINED_INFO)})"
        raise ValueError(msg)
    info = PRETRAINED_INFO[pretrained_model]
    mod = importlib.import_module(f"tiatoolbox_b200.models.architecture.{info['class'][0]}")
    model = getattr(mod, info["class"][1])(**info["kwargs"])
    if pretrained_weights is not None:
        model.load_weights_from_file(pretrained_weights)
    else:
        logger.warning("no weights given for `%s` and no network access: the model keeps seeded random weights", pretrained_model)
    return model, info["ioconfig"]
