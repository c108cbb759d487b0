"""tiab_net_create / tiab_net_infer: a serialised device plan run through the C ABI alone.

1. Python drives the three C calls through ctypes (`export.NetRunner`: no torch op between input and outputs except buffer allocation)
   and the outputs must equal `plan.forward` bit for bit — same kernels, same arguments, a different host.
2. A C program (tests/c/net_driver.c, compiled here with gcc against include/tiab200.h + libcudart, no Python, no torch) loads the plan
   FILE, feeds a raw uint8 tile file and writes the heads to disk: what a cgo / JNI / Rust binder would do.  Its outputs must equal
   the Python plan's as well.
"""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.cnn_common import hovernet_test_weights, resnet18_test_weights, unet_test_weights

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _plans():
    from tiatoolbox_b200.models.hovernet_plan import HoVerNetPlan
    from tiatoolbox_b200.models.resnet_plan import CNNPlan, UNetPlan

    yield "hovernet_fast", HoVerNetPlan(hovernet_test_weights("fast", 6), batch=2, device="cuda:0", mode="fast", num_types=6), 256
    yield "resnet18", CNNPlan(resnet18_test_weights(9), batch=4, device="cuda:0", backbone="resnet18", num_classes=9, in_size=224), 224
    yield "unet", UNetPlan(unet_test_weights(), batch=1, device="cuda:0", in_size=256), 256


def test_exported_plans_run_through_the_c_abi(built_lib, tmp_path):
    import torch

    from tiatoolbox_b200.models.export import NetRunner, export_plan
    from tiatoolbox_b200.synthetic import he_tiles

    for name, plan, size in _plans():
        plan.capture()
        b = plan.inp.shape[0]
        x = torch.from_numpy(he_tiles(b, size, size, seed=3)).cuda()
        want = plan.forward(x)
        want = [w.clone() for w in (want if isinstance(want, (tuple, list)) else [want])]
        blob = export_plan(plan)
        net = NetRunner(blob)
        assert (net.batch, net.in_h, net.in_w, net.in_c, net.n_outputs) == (b, size, size, 3, len(want))
        got = net.infer(x)
        torch.cuda.synchronize()
        for g, w in zip(got, want):
            assert g.shape == w.shape and torch.equal(g, w), name
        # a partial batch: the first rows are the first images' results
        if b > 1:
            got1 = net.infer(x[:1].contiguous())
            torch.cuda.synchronize()
            assert all(torch.equal(g, w[:1]) for g, w in zip(got1, want)), name
        if name == "hovernet_fast":   # the same plan file through a C program
            gcc, cuda = shutil.which("gcc"), Path("/usr/local/cuda")
            if gcc is None or not (cuda / "include" / "cuda_runtime_api.h").exists():
                pytest.skip("no gcc / CUDA headers for the C driver")
            exe = tmp_path / "net_driver"
            subprocess.run([gcc, "-O1", "-o", str(exe), str(ROOT / "tests" / "c" / "net_driver.c"), f"-I{ROOT / 'include'}", f"-I{cuda / 'include'}",
                            f"-L{ROOT / 'tiatoolbox_b200'}", f"-L{cuda / 'lib64'}", "-ltiab200", "-lcudart",
                            f"-Wl,-rpath,{ROOT / 'tiatoolbox_b200'}", f"-Wl,-rpath,{cuda / 'lib64'}"], check=True)
            (tmp_path / "plan.bin").write_bytes(blob)
            x.cpu().numpy().tofile(tmp_path / "tiles.u8")
            r = subprocess.run([str(exe), str(tmp_path / "plan.bin"), str(tmp_path / "tiles.u8"), str(tmp_path / "head")], capture_output=True, text=True, check=False)
            assert r.returncode == 0, r.stdout + r.stderr
            for k, w in enumerate(want):
                got_c = np.fromfile(tmp_path / f"head{k}.f32", dtype=np.float32).reshape(tuple(w.shape))
                assert np.array_equal(got_c, w.cpu().numpy()), (name, k)
        net.close()
        del plan
        torch.cuda.empty_cache()


def test_net_create_rejects_garbage(built_lib):
    import ctypes as C

    from tiatoolbox_b200 import _lib

    h = _lib.handle(0)
    net = C.c_void_p()
    junk = (C.c_char * 64)(*b"not a plan")
    assert _lib.lib().tiab_net_create(h, junk, 64, C.byref(net)) != 0 and not net
    assert b"TIABPLN1" in _lib.lib().tiab_last_error(h)
