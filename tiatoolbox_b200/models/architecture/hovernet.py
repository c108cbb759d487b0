"""HoVer-Net (Graham et al. 2019) for the B200 — architecture definition + fp32 reference forward.

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
import torch.nn.functional as F

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


def _bn(x, sd, name):
    if sd.get("__calibrate__", False):  # test-weight calibration: adopt the batch statistics (see _params.calibrate)
        sd[name + ".running_mean"] = x.mean(dim=(0, 2, 3)).detach()
        sd[name + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).detach() + 1e-3
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.0, 1e-5)


def _crop_to(x, ref):
    dh, dw = x.shape[2] - ref.shape[2], x.shape[3] - ref.shape[3]
    return x[:, :, dh // 2: x.shape[2] - (dh - dh // 2), dw // 2: x.shape[3] - (dw - dw // 2)]


def _up2(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _residual_block(x, sd, name, units, stride):
    shortcut = F.conv2d(x, sd[f"{name}.shortcut.weight"], stride=stride) if f"{name}.shortcut.weight" in sd else x
    prev = x
    for i in range(units):
        u = f"{name}.units.{i}"
        y = prev
        if i > 0:
            y = F.relu(_bn(y, sd, f"{u}.preact/bn"))
        y = F.relu(_bn(F.conv2d(y, sd[f"{u}.conv1.weight"]), sd, f"{u}.conv1/bn"))
        s = stride if i == 0 else 1
        pb, pa = tf_same_pad(y.shape[2], 3, s)
        y = F.pad(y, (pb, pa, pb, pa))
        y = F.relu(_bn(F.conv2d(y, sd[f"{u}.conv2.weight"], stride=s), sd, f"{u}.conv2/bn"))
        y = F.conv2d(y, sd[f"{u}.conv3.weight"])
        prev = y + shortcut
        shortcut = prev
    return F.relu(_bn(prev, sd, f"{name}.blk_bna.bn"))


def _dense_block(x, sd, name, units):
    prev = x
    for i in range(units):
        u = f"{name}.units.{i}"
        y = F.relu(_bn(prev, sd, f"{u}.preact_bna/bn"))
        y = F.relu(_bn(F.conv2d(y, sd[f"{u}.conv1.weight"]), sd, f"{u}.conv1/bn"))
        y = F.conv2d(y, sd[f"{u}.conv2.weight"], groups=4)
        prev = torch.cat([_crop_to(prev, y), y], dim=1)
    return F.relu(_bn(prev, sd, f"{name}.blk_bna.bn"))


def hovernet_forward_reference(sd, x_nchw: torch.Tensor, mode: str, num_types: int | None, capture: dict | None = None,
                               num_layers: int | None = None) -> OrderedDict:
    """fp32 PyTorch restatement of HoVerNet.forward (hovernet.py:405-454; with `num_layers` HoVerNetPlus, which only adds the `ls`
    decoder branch: hovernetplus.py:87-133).  x in [0, 255]."""
    x = x_nchw / 255.0
    if mode == "fast":
        pb, pa = tf_same_pad(x.shape[2], 7, 1)
        x = F.pad(x, (pb, pa, pb, pa))
    d = F.relu(_bn(F.conv2d(x, sd["conv0./.weight"]), sd, "conv0.bn"))
    if capture is not None:
        capture["x0"] = d
    feats = []
    for name, _, _, units, stride in RES_CFG:
        d = _residual_block(d, sd, name, units, stride)
        feats.append(d)
        if capture is not None:
            capture[name] = d
    feats[3] = F.conv2d(feats[3], sd["conv_bot.weight"])
    if capture is not None:
        capture["dbot"] = feats[3]
    c0, c1 = (184, 72) if mode == "original" else (92, 36)
    d0 = feats[0][:, :, c0 // 2: feats[0].shape[2] - c0 // 2, c0 // 2: feats[0].shape[3] - c0 // 2]
    d1 = feats[1][:, :, c1 // 2: feats[1].shape[2] - c1 // 2, c1 // 2: feats[1].shape[3] - c1 // 2]
    k = 5 if mode == "original" else 3
    out = OrderedDict()
    for b in branch_channels(num_types, num_layers):
        p = f"decoder.{b}"
        u3 = _up2(feats[3]) + feats[2]
        if capture is not None:
            capture["u3in"] = u3
        u3 = F.conv2d(u3, sd[f"{p}.u3.conva.weight"])
        if capture is not None:
            capture[f"{b}.u3.conva"] = u3
        u3 = F.conv2d(_dense_block(u3, sd, f"{p}.u3.dense", 8), sd[f"{p}.u3.convf.weight"])
        if capture is not None:
            capture[f"{b}.u3out"] = u3
        u2 = _up2(u3) + d1
        if capture is not None:
            capture[f"{b}.u2in"] = u2
        u2 = F.conv2d(u2, sd[f"{p}.u2.conva.weight"])
        u2 = F.conv2d(_dense_block(u2, sd, f"{p}.u2.dense", 4), sd[f"{p}.u2.convf.weight"])
        if capture is not None:
            capture[f"{b}.u2out"] = u2
        u1 = _up2(u2) + d0
        if capture is not None:
            capture[f"{b}.u1in"] = u1
        pb, pa = tf_same_pad(u1.shape[2], k, 1)
        u1 = F.conv2d(F.pad(u1, (pb, pa, pb, pa)), sd[f"{p}.u1.conva.weight"])
        if capture is not None:
            capture[f"{b}.u1"] = u1
        u0 = F.relu(_bn(u1, sd, f"{p}.u0.bn"))
        out[b] = F.conv2d(u0, sd[f"{p}.u0.conv.weight"], sd[f"{p}.u0.conv.bias"])
    return out


def hovernet_heads_reference(out: OrderedDict) -> tuple[torch.Tensor, ...]:
    """Head activations of HoVerNet.infer_batch (hovernet.py:897-909): NHWC, np = softmax[..., 1:],
    tp = argmax as float32, hv raw.  Returns (np, hv[, tp])."""
    o = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in out.items()}
    o["np"] = F.softmax(o["np"], dim=-1)[..., 1:]
    if "tp" in o:
        tp = torch.argmax(F.softmax(o["tp"], dim=-1), dim=-1, keepdim=True).type(torch.float32)
        if "ls" in o:   # HoVerNetPlus.infer_batch (hovernetplus.py:444-459)
            ls = torch.argmax(F.softmax(o["ls"], dim=-1), dim=-1, keepdim=True).type(torch.float32)
            return o["np"], o["hv"], tp, ls
        return o["np"], o["hv"], tp
    return o["np"], o["hv"]


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
        """fp32 PyTorch forward (NCHW in [0, 255]) — the parity reference, not the hot path."""
        sd = {k: v.to(input_tensor.device) for k, v in self._sd.items()}
        return hovernet_forward_reference(sd, input_tensor, self.mode, self.num_types)

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
