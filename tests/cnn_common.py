"""Shared helpers of the CNN tests: seeded + calibrated test weights (bit-identical on every box)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def hovernet_test_weights(mode="fast", num_types=6):
    from tiatoolbox_b200.models.architecture._params import seeded_state_dict
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec

    sd = seeded_state_dict(hovernet_spec(3, num_types, mode), seed=0)
    calib = np.load(GOLDEN / f"hovernet_{mode}_bn_calib.npz")
    for k in calib.files:
        sd[k] = torch.from_numpy(calib[k])
    return sd


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _with_calib(sd, fname):
    calib = np.load(GOLDEN / fname)
    for k in calib.files:
        sd[k] = torch.from_numpy(calib[k])
    return sd


def resnet18_test_weights(num_classes=9):
    from tiatoolbox_b200.models.architecture._params import seeded_state_dict
    from tiatoolbox_b200.models.architecture.vanilla import cnn_spec

    return _with_calib(seeded_state_dict(cnn_spec("resnet18", num_classes), seed=0), "resnet18_bn_calib.npz")


def unet_test_weights():
    from tiatoolbox_b200.models.architecture._params import seeded_state_dict
    from tiatoolbox_b200.models.architecture.unet import unet_spec

    return _with_calib(seeded_state_dict(unet_spec(3, 2, (3,)), seed=0), "unet_bn_calib.npz")
