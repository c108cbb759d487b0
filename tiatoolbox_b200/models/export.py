"""Serialised device plans: `export_plan(plan)` turns a built network plan (HoVerNetPlan / CNNPlan / UNetPlan: a fixed list of
libtiab200 launches over pre-allocated buffers) into ONE self-contained binary blob that `tiab_net_create` (include/tiab200.h, csrc/net.cu)
loads without any Python: buffer table, constant data (packed weights, folded BatchNorm vectors), and the launch list with every pointer
replaced by (buffer id, byte offset).  A non-Python binder runs a model with three C calls: tiab_net_create(blob) -> tiab_net_infer(...) ->
tiab_net_destroy — the deployable artefact is the plan file, like an engine file of an inference runtime.

The plan is recorded, not re-derived: the exporter runs the plan's launch list once through a recording proxy of the C library, so the
blob holds exactly the calls `plan.forward` makes (the reference-side meaning of the model stays in the Python plan builders:
models/hovernet_plan.py, models/resnet_plan.py — ModelABC.infer_batch, models_abc.py:105-131).

Blob layout (little endian):
  header   : magic "TIABPLN1" | u32 n_buffers | u32 n_ops | u32 batch | u32 in_h | u32 in_w | u32 in_c | u32 input_buffer | u32 n_outputs
  outputs  : n_outputs x { u32 buffer | u64 byte offset | u32 ndim | u32 dims[4] | u32 dtype (0 = f32) }
  buffers  : n_buffers x { u64 nbytes | u32 kind (0 = zero-initialised scratch, 1 = constant) | u64 data offset (kind 1) }
  ops      : n_ops x { u32 function id | u32 n_args | args: { u32 type | u64 a | u64 b } }
             type 0 int (a), 1 double (a = bits), 2 pointer (a = buffer id, b = byte offset), 3 null pointer,
             4 conv descriptor (a = byte offset of the descriptor block, b = its length): the raw tiab_conv_desc_t followed by
               u32 n_reloc x { u32 field byte offset | u32 buffer id | u64 byte offset }
  data     : constant payloads and descriptor blocks, 16-byte aligned
"""
from __future__ import annotations

import ctypes as C
import struct

import torch

from tiatoolbox_b200 import _lib
from tiatoolbox_b200.models.ops import Act

FUNC_IDS = {"tiab_conv_gemm": 1, "tiab_im2col": 2, "tiab_bn_relu": 3, "tiab_upsample_add": 4, "tiab_maxpool3x3s2": 5,
            "tiab_pointwise_head": 6, "tiab_gap_fc_softmax": 7, "tiab_hovernet_heads": 8, "tiab_softmax_up2_crop": 9}
MAGIC = b"TIABPLN1"


def _tensors_in(obj, out: list, depth: int = 0) -> None:
    if depth > 4 or obj is None:
        return
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            out.append(obj)
        return
    if isinstance(obj, Act):
        out.append(obj.t)
        return
    if isinstance(obj, (list, tuple)):
        for o in obj:
            _tensors_in(o, out, depth + 1)
    elif isinstance(obj, dict):
        for o in obj.values():
            _tensors_in(o, out, depth + 1)


class _Recorder:
    """Stands in for the ctypes library while the plan's launch list runs once: logs (name, args), then makes the real call."""

    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if name not in FUNC_IDS:
            return fn

        def call(*args):
            self.calls.append((name, args))
            return fn(*args)

        return call


def plan_outputs(plan) -> list[torch.Tensor]:
    """The tensors `plan.forward` returns (device buffers of the plan)."""
    if getattr(plan, "np_out", None) is not None:
        outs = [plan.np_out, plan.hv_out]
        if getattr(plan, "tp_out", None) is not None:
            outs.append(plan.tp_out)
        return outs
    return [plan.probs]


def export_plan(plan) -> bytes:
    if getattr(plan, "ls_out", None) is not None:
        msg = "HoVerNetPlus applies a torch softmax / argmax after the plan (layer head): export the three-branch models"
        raise NotImplementedError(msg)
    # ---- every device tensor the launches may touch ----
    cands: list[torch.Tensor] = []
    _tensors_in(list(plan.bufs.values()), cands)
    _tensors_in(plan._keep, cands)  # noqa: SLF001
    for op in plan.ops:
        for attr in ("_keep", "_oscale_keep", "_head_keep"):
            _tensors_in(getattr(op, attr, None), cands)
        if getattr(op, "__closure__", None):
            _tensors_in([c.cell_contents for c in op.__closure__ if not callable(c.cell_contents)], cands)
        _tensors_in(getattr(op, "__defaults__", None), cands)
    storages = {}
    for t in cands:
        st = t.untyped_storage()
        storages.setdefault(st.data_ptr(), (st, st.nbytes()))
    bases = sorted(storages)
    scratch = {t.untyped_storage().data_ptr() for t in plan.bufs.values()}
    buf_id = {b: i for i, b in enumerate(bases)}

    def locate(ptr: int):
        import bisect

        i = bisect.bisect_right(bases, ptr) - 1
        if i < 0 or ptr >= bases[i] + storages[bases[i]][1]:
            raise RuntimeError(f"pointer {ptr:#x} of a recorded launch lies in no tensor of the plan")
        return buf_id[bases[i]], ptr - bases[i]

    # ---- record one eager pass ----
    real = _lib.lib()
    rec = _Recorder(real)
    orig = _lib.lib
    _lib.lib = lambda: rec
    try:
        plan.run_eager()
        torch.cuda.synchronize(plan.dev)
    finally:
        _lib.lib = orig

    data = bytearray()

    def put(blob: bytes) -> int:
        while len(data) % 16:
            data.append(0)
        off = len(data)
        data.extend(blob)
        return off

    ptr_fields = []
    for name, typ in _lib.ConvDesc._fields_:  # noqa: SLF001
        off = getattr(_lib.ConvDesc, name).offset
        if typ is C.c_void_p:
            ptr_fields.append((off, None))
        elif isinstance(typ, type) and issubclass(typ, C.Array) and typ._type_ is C.c_void_p:  # noqa: SLF001
            ptr_fields += [(off + 8 * k, None) for k in range(typ._length_)]  # noqa: SLF001

    ops_bin = bytearray()
    for name, args in rec.calls:
        sig = _lib.SIGNATURES[name][1]
        assert len(sig) == len(args), name
        enc = []
        for typ, a in list(zip(sig, args))[1:-1]:          # first = handle, last = stream: supplied by the C runner
            if typ is _lib._vp:  # noqa: SLF001
                v = a.value if isinstance(a, C.c_void_p) else a
                if v is None or v == 0:
                    enc.append((3, 0, 0))
                else:
                    b, o = locate(int(v))
                    enc.append((2, b, o))
            elif isinstance(typ, type) and issubclass(typ, C._Pointer):  # noqa: SLF001 - byref(ConvDesc)
                desc = a._obj  # noqa: SLF001
                raw = bytes(desc)
                relocs = []
                for off, _ in ptr_fields:
                    v = struct.unpack_from("<Q", raw, off)[0]
                    if v:
                        b, o = locate(v)
                        relocs.append((off, b, o))
                block = raw + struct.pack("<I", len(relocs)) + b"".join(struct.pack("<IIQ", *r) for r in relocs)
                enc.append((4, put(block), len(raw)))
            elif typ in (C.c_float, C.c_double):
                enc.append((1, struct.unpack("<Q", struct.pack("<d", float(a)))[0], 0))
            else:
                enc.append((0, int(a) & 0xFFFFFFFFFFFFFFFF, 0))
        ops_bin += struct.pack("<II", FUNC_IDS[name], len(enc)) + b"".join(struct.pack("<IQQ", *e) for e in enc)

    bufs_bin = bytearray()
    for b in bases:
        st, nbytes = storages[b]
        if b in scratch:
            bufs_bin += struct.pack("<QIQ", nbytes, 0, 0)
        else:
            host = torch.empty(nbytes, dtype=torch.uint8)
            host.untyped_storage().copy_(st)
            bufs_bin += struct.pack("<QIQ", nbytes, 1, put(host.numpy().tobytes()))

    outs = plan_outputs(plan)
    outs_bin = bytearray()
    for t in outs:
        assert t.dtype == torch.float32 and t.is_contiguous()
        b, o = locate(t.data_ptr())
        dims = list(t.shape) + [1] * (4 - t.dim())
        outs_bin += struct.pack("<IQI4II", b, o, t.dim(), *dims[:4], 0)
    in_b, in_o = locate(plan.inp.data_ptr())
    assert in_o == 0
    hdr = MAGIC + struct.pack("<8I", len(bases), len(rec.calls), plan.inp.shape[0], plan.inp.shape[1], plan.inp.shape[2], plan.inp.shape[3],
                              in_b, len(outs))
    head = hdr + bytes(outs_bin) + bytes(bufs_bin) + bytes(ops_bin)
    pad = (-len(head)) % 16
    # data offsets are relative to the start of the data section
    return head + b"\0" * pad + struct.pack("<Q", len(data)) + b"\0" * 8 + bytes(data)


class NetRunner:
    """Thin ctypes driver of the C network runner (tiab_net_create / tiab_net_infer / tiab_net_destroy): what a non-Python binder does."""

    def __init__(self, blob: bytes, device: int = 0):
        self._lib = _lib.lib()
        self._h = _lib.handle(device)
        self._net = C.c_void_p()
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(self._h, self._lib.tiab_net_create(self._h, buf, len(blob), C.byref(self._net)))
        info = (C.c_int * 64)()
        _lib.check(self._h, self._lib.tiab_net_info(self._net, info, 64))
        self.batch, self.in_h, self.in_w, self.in_c, self.n_outputs = (int(info[i]) for i in range(5))
        self.out_shapes = [tuple(int(info[5 + 5 * k + 1 + d]) for d in range(int(info[5 + 5 * k]))) for k in range(self.n_outputs)]

    def infer(self, x_u8: torch.Tensor) -> list[torch.Tensor]:
        n = x_u8.shape[0]
        assert x_u8.is_cuda and x_u8.dtype == torch.uint8 and x_u8.is_contiguous() and tuple(x_u8.shape[1:]) == (self.in_h, self.in_w, self.in_c)
        outs = [torch.empty((n, *s[1:]), dtype=torch.float32, device=x_u8.device) for s in self.out_shapes]
        ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        _lib.check(self._h, self._lib.tiab_net_infer(self._net, x_u8.data_ptr(), n, ptrs, torch.cuda.current_stream(x_u8.device).cuda_stream))
        return outs

    def close(self):
        if self._net:
            self._lib.tiab_net_destroy(self._net)
            self._net = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001, S110
            pass


def main(argv=None) -> int:
    """`python -m tiatoolbox_b200.models.export <registry name> <out.plan> [--weights w.pth] [--batch 16] [--size 256] [--device 0]`
    writes the plan file a non-Python host loads with tiab_net_create.  Registry names: architecture/__init__.py."""
    import argparse
    from pathlib import Path

    from tiatoolbox_b200.models.architecture import get_pretrained_model

    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("model")
    ap.add_argument("out")
    ap.add_argument("--weights", default=None, help="tiatoolbox checkpoint (.pth); default: the seeded random weights")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=None, help="input tile side (default: the registry ioconfig's patch_input_shape)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    model, io = get_pretrained_model(a.model, a.weights)
    dev = torch.device("cuda", a.device)
    model.to(dev)
    size = a.size or int(io.patch_input_shape[0])
    plan = model.plan_for(a.batch, dev, (size, size))
    blob = export_plan(plan)
    Path(a.out).write_bytes(blob)
    print(f"{a.model}: batch {a.batch}, {size}x{size} tiles, {len(plan.ops)} launches, {len(blob) / 1e6:.1f} MB -> {a.out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
