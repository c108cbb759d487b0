"""Parameter bookkeeping shared by the architecture definitions.

The B200 models keep their weights in a flat ``{name: tensor}`` dict whose keys are exactly the
reference modules' ``state_dict`` keys, so checkpoints written by / for tiatoolbox load unchanged
(``load_state_dict``), and build two things from it: the fp32 PyTorch reference forward
(``forward``) and the fused fp16 device plan.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import torch


class ParamSpec(OrderedDict):
    """name -> (shape, kind) with kind in {conv, bn_w, bn_b, bn_mean, bn_var, bn_n, fc_w, bias}."""

    def conv(self, name, cout, cin, k, bias=False, gain=1.0):
        self[name + ".weight"] = ((cout, cin, k, k), "conv" if gain == 1.0 else f"conv:{gain}")
        if bias:
            self[name + ".bias"] = ((cout,), "bias")

    def bn(self, name, c, gain=1.0):
        self[name + ".weight"] = ((c,), "bn_w" if gain == 1.0 else f"bn_w:{gain}")
        self[name + ".bias"] = ((c,), "bn_b")
        self[name + ".running_mean"] = ((c,), "bn_mean")
        self[name + ".running_var"] = ((c,), "bn_var")
        self[name + ".num_batches_tracked"] = ((), "bn_n")

    def linear(self, name, cout, cin):
        self[name + ".weight"] = ((cout, cin), "fc_w")
        self[name + ".bias"] = ((cout,), "bias")


def seeded_state_dict(spec: ParamSpec, seed: int = 0) -> OrderedDict:
    """Deterministic random weights: every tensor is drawn from its own generator keyed by
    (seed, name), so the values do not depend on construction order or on torch's module code.
    He-normal convolutions; BatchNorm affine and running statistics are randomised too so that
    BN folding is exercised."""
    sd = OrderedDict()
    for name, (shape, kind) in spec.items():
        key = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:6], "little")
        g = torch.Generator().manual_seed(key)
        if kind.startswith("conv"):
            # "conv:<gain>": the closing conv of a residual branch gets a small gain, as in trained ResNets
            # (zero-init-residual / small last-BN gamma); with gain 1 a random residual network is chaotic and
            # amplifies fp32 round-off differences (summation order!) by ~5x per block.
            gain = float(kind.split(":")[1]) if ":" in kind else 1.0
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5 * gain
        elif kind == "fc_w":
            t = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif kind.startswith("bn_w"):
            # "bn_w:<gain>": small gamma on the BN that closes a residual branch (zero-init-residual practice)
            gain = float(kind.split(":")[1]) if ":" in kind else 1.0
            t = (torch.rand(shape, generator=g) * 0.4 + 0.8) * gain
        elif kind in ("bn_b", "bias", "bn_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif kind == "ones":
            t = torch.ones(shape)
        else:  # num_batches_tracked
            t = torch.zeros(shape, dtype=torch.long)
        sd[name] = t
    return sd


def bn_scale_shift(sd, name, eps=1e-5):
    """Eval-mode BatchNorm as y = x*scale + shift (float64 fold, float32 result)."""
    w = sd[name + ".weight"].double()
    b = sd[name + ".bias"].double()
    m = sd[name + ".running_mean"].double()
    v = sd[name + ".running_var"].double()
    scale = w / torch.sqrt(v + eps)
    return scale.float(), (b - m * scale).float()


def calibrate(sd, forward_fn, x):
    """Give seeded random weights realistic BatchNorm running statistics.

    Untrained BN statistics let a pre-activation residual stream double in variance per unit
    (activations reach 1e4 after 16 units) — nothing like a trained network and hostile to fp16.
    One forward pass in which every BN adopts the statistics of its input (as training would have
    converged to) yields O(1) activations everywhere.  Returns {name: tensor} of the adopted
    running_mean / running_var entries (committed as a small fixture so that every machine builds
    bit-identical test weights)."""
    sd["__calibrate__"] = True
    try:
        with torch.inference_mode():
            forward_fn(sd, x)
    finally:
        del sd["__calibrate__"]
    return {k: v.clone() for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}


def with_calibration(sd, npz_path):
    """Overlay committed BatchNorm running statistics (see `calibrate`) on a seeded state dict."""
    import numpy as np

    calib = np.load(npz_path)
    for k in calib.files:
        sd[k] = torch.from_numpy(calib[k])
    return sd
