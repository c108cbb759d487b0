"""HoVer-Net (Graham et al. 2019) for the B200 — architecture definition (parameter names / shapes) and the model class.

Mirrors tiatoolbox/models/architecture/hovernet.py: same constructor arguments, same state_dict
keys (so `hovernet_fast-pannuke` / `hovernet_original-*` checkpoints load unchanged), same
`infer_batch` / `postproc` contract.  Layer structure restated from hovernet.py:30-262 (blocks) and
:330-500 (network + decoder branches):

  conv0: [TF-same pad (fast)] 7x7 s1 -> BN -> ReLU
  d0..d3: pre-activation bottleneck residual blocks (3,4,6,3 units; 1x1 -> 3x3 (TF-same pad, stride on
          unit 0) -> 1x1; 1x1 shortcut conv on unit 0), closing BN+ReLU
  conv_bot 1x1 2048->1024
  per branch (tp?, np, hv): u3 = up2x(d3)+crop(d2) -> conv k valid -> Dense(8 units, g=4) -> 1x1
                            u2 = up2x(u3)+crop(d1) -> conv k valid -> Dense(4 units)      -> 1x1
                            u1 = up2x(u2)+crop(d0) -> TF-same pad -> conv k ; u0 = BN-ReLU-1x1(+bias)
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from tiatoolbox_b200.models.architecture._params import ParamSpec, seeded_state_dict

RES_CFG = (("d0", 64, (64, 64, 256), 3, 1), ("d1", 256, (128, 128, 512), 4, 2),
           ("d2", 512, (256, 256, 1024), 6, 2), ("d3", 1024, (512, 512, 2048), 3, 2))


def branch_channels(num_types: int | None, num_layers: int | None = None) -> OrderedDict:
    """Decoder branches in the reference's ModuleDict order with their output channels: [tp], np, hv (hovernet.py:381-403) and, for
    HoVerNet+, ls (hovernetplus.py:107-131)."""
    br = OrderedDict()
    if num_types is not None:
        br["tp"] = num_types
    br["np"] = 2
    br["hv"] = 2
    if num_layers is not None:
        br["ls"] = num_layers
    return br


def hovernet_spec(num_input_channels: int = 3, num_types: int | None = None, mode: str = "original",
                  num_layers: int | None = None) -> ParamSpec:
    sp = ParamSpec()
    sp.conv("conv0./", 64, num_input_channels, 7)
    sp.bn("conv0.bn", 64)
    for name, cin, (c1, c2, c3), units, stride in RES_CFG:
        ch = cin
        for i in range(units):
            u = f"{name}.units.{i}"
            if i > 0:
                sp.bn(f"{u}.preact/bn", ch)
            sp.conv(f"{u}.conv1", c1, ch, 1)
            sp.bn(f"{u}.conv1/bn", c1)
            sp.conv(f"{u}.conv2", c2, c1, 3)
            sp.bn(f"{u}.conv2/bn", c2)
            sp.conv(f"{u}.conv3", c3, c2, 1, gain=0.2)
            ch = c3
        if cin != c3 or stride != 1:
            sp.conv(f"{name}.shortcut", c3, cin, 1)
        sp.bn(f"{name}.blk_bna.bn", c3)
    sp.conv("conv_bot", 1024, 2048, 1)
    k = 5 if mode == "original" else 3
    for b, out_ch in branch_channels(num_types, num_layers).items():
        for lvl, cin, cmid, units in (("u3", 1024, 256, 8), ("u2", 512, 128, 4)):
            p = f"decoder.{b}.{lvl}"
            sp.conv(f"{p}.conva", cmid, cin, k)
            ch = cmid
            for i in range(units):
                u = f"{p}.dense.units.{i}"
                sp.bn(f"{u}.preact_bna/bn", ch)
                sp.conv(f"{u}.conv1", 128, ch, 1)
                sp.bn(f"{u}.conv1/bn", 128)
                sp[f"{u}.conv2.weight"] = ((32, 128 // 4, k, k), "conv")  # groups = 4
                ch += 32
            sp.bn(f"{p}.dense.blk_bna.bn", ch)
            sp.conv(f"{p}.convf", ch, ch, 1)
        sp.conv(f"decoder.{b}.u1.conva", 64, 256, k)
        sp.bn(f"decoder.{b}.u0.bn", 64)
        sp.conv(f"decoder.{b}.u0.conv", out_ch, 64, 1, bias=True)
    sp["upsample2x.unpool_mat"] = ((2, 2), "ones")  # buffer of the reference's UpSample2x (utils.py:202-243)
    return sp


def tf_same_pad(size: int, ksize: int, stride: int) -> tuple[int, int]:
    """(before, after) zero padding of TFSamepaddingLayer (hovernet.py:44-58)."""
    pad = max(ksize - stride, 0) if size % stride == 0 else max(ksize - (size % stride), 0)
    return pad // 2, pad - pad // 2


# =================================================================================================
# Model class (API mirror of tiatoolbox/models/architecture/hovernet.py:264-910)
# =================================================================================================
from tiatoolbox_b200.models.models_abc import ModelABC  # noqa: E402


class HoVerNet(ModelABC):
    """HoVer-Net on the B200.  Same constructor / `infer_batch` / `postproc` contract as the reference."""

    def __init__(self, num_input_channels: int = 3, num_types: int | None = None, mode: str = "original",
                 nuc_type_dict: dict | None = None) -> None:
        super().__init__()
        if mode not in ["original", "fast"]:
            msg = f"Invalid mode {mode} for HoVerNet. Only support `original` or `fast`."
            raise ValueError(msg)
        self.mode = mode
        self.num_types = num_types
        self.nuc_type_dict = nuc_type_dict
        self.tasks = ["nuclei_segmentation"]
        self.class_dict = {self.tasks[0]: nuc_type_dict}
        self.num_input_channels = num_input_channels
        self._sd = seeded_state_dict(hovernet_spec(num_input_channels, num_types, mode), seed=0)

    def forward(self, input_tensor: torch.Tensor) -> OrderedDict:
        """hovernet.py:405-454 on the device plan: NCHW input holding uint8 pixel values in [0, 255] -> OrderedDict of NCHW fp32 logits
        per decoder branch.  There is no PyTorch-eager forward in the product (the fp32 restatement is test infrastructure: oracle/)."""
        x = ModelABC._nchw_to_u8_nhwc(input_tensor, self._device)
        plan = self.plan_for(x.shape[0], self._device, x.shape[1:3])
        plan.forward(x)
        return OrderedDict((b, plan.logits[b].permute(0, 3, 1, 2).clone()) for b in plan.logits)

    def _build_plan(self, batch, device, shape):
        from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan

        size = None if shape is None else int(shape[0])
        return HoVerNetPlan(self._sd, batch=batch, device=device, mode=self.mode, num_types=self.num_types, in_size=size,
                            precision=self.precision)

    @staticmethod
    def infer_batch_device(model: "HoVerNet", batch_data, *, device) -> tuple[torch.Tensor, ...]:
        """Like infer_batch but the head maps stay on the device (views valid until the next call)."""
        x = ModelABC._as_device_u8(batch_data, device)
        n = x.shape[0]
        plan_b = getattr(model, "_plan_batch", None) or n
        if n > plan_b:
            msg = f"batch of {n} exceeds the plan batch {plan_b}"
            raise ValueError(msg)
        plan = model.plan_for(plan_b, device, x.shape[1:3])
        if n < plan_b:
            pad = torch.zeros((plan_b - n, *x.shape[1:]), dtype=torch.uint8, device=x.device)
            x = torch.cat([x, pad])
        outs = plan.forward(x)
        return tuple(o[:n] for o in outs)

    @staticmethod
    def infer_batch(model: "HoVerNet", batch_data, *, device) -> tuple[np.ndarray, ...]:
        """hovernet.py:861-910: (np, hv[, tp]) float32 numpy arrays, NHWC."""
        outs = HoVerNet.infer_batch_device(model, batch_data, device=device)
        return tuple(o.cpu().numpy() for o in outs)

    def postproc(self, raw_maps, offset: tuple[int, int] = (0, 0)) -> tuple[dict, ...]:
        """hovernet.py:751-859 for one image: device watershed + instance table, host contours."""
        from tiatoolbox_b200.models import postproc as pp

        dev = self._device
        if len(raw_maps) == 3:
            np_map, hv_map, tp_map = raw_maps
        else:
            np_map, hv_map = raw_maps
            tp_map = None
        t = [torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev)[None] for a in (np_map, hv_map)]
        tt = None if tp_map is None else torch.as_tensor(np.asarray(tp_map), dtype=torch.float32).to(dev)[None]
        inst, table = pp.proc_np_hv_batch(t[0], t[1], tt)
        pred_inst = inst[0].cpu().numpy()
        row = {k: v[0].cpu().numpy() for k, v in table.items()}
        info = pp.instance_info(pred_inst, row, offset)
        return (pack_instance_result(self.tasks[0], pred_inst, info),)


def pack_instance_result(task: str, pred_inst: np.ndarray, info: dict) -> dict:
    """The reference's per-image result dict (hovernet.py:828-859 + _inst_dict_for_dask_processing :913-932)."""
    if not info:
        empty = np.empty(shape=0)
        info_ = {"box": empty, "centroid": empty, "contours": empty, "prob": empty, "type": empty}
    else:
        info_ = {}
        for key in ("box", "centroid", "contours", "prob", "type"):
            col = np.empty(len(info), dtype=object)
            for i, rec in enumerate(info.values()):
                col[i] = rec[key]
            info_[key] = col
    return {"task_type": task, "predictions": pred_inst, "info_dict": info_, "seg_type": "instance"}
