"""GPU parity of the resnet18 classifier plan and the ResNet-50 U-Net plan against the plain-PyTorch fp32
restatement (same GPU, TF32 off) and against the unmodified reference modules' CPU outputs (tests/golden).
Tolerance: probabilities within 1e-3 relative (relative L2) and max|delta| <= 1e-3; argmax identical."""
from __future__ import annotations

import numpy as np
import pytest

from tests.cnn_common import GOLDEN, rel_l2, resnet18_test_weights, unet_test_weights

pytestmark = pytest.mark.gpu


def _no_tf32():
    import torch

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_resnet18_probabilities(built_lib):
    import torch

    from oracle.torch_reference import cnn_forward_reference
    from tiatoolbox_b200.models.resnet_plan import CNNPlan
    from tiatoolbox_b200.synthetic import he_tiles

    _no_tf32()
    sd = resnet18_test_weights()
    tiles = he_tiles(16, 224, 224, seed=0)
    x = torch.from_numpy(tiles).cuda()
    gold = np.load(GOLDEN / "resnet18_golden.npz")["probs"]
    with torch.inference_mode():
        ref = cnn_forward_reference({k: v.cuda() for k, v in sd.items()}, x.permute(0, 3, 1, 2).float() / 255.0, "resnet18").cpu().numpy()
    for prec, tol in (("fp16x3", 1e-3), ("fp16", 2e-2)):
        plan = CNNPlan(sd, backbone="resnet18", num_classes=9, batch=16, device="cuda:0", precision=prec)
        probs = plan.forward(x).cpu().numpy()
        assert probs.shape == (16, 9) and np.allclose(probs.sum(1), 1, atol=1e-5)
        e1, e2 = rel_l2(probs, ref), rel_l2(probs, gold)
        print(f"resnet18 {prec}: rel_l2 vs torch fp32 {e1:.2e}, vs reference golden {e2:.2e}")
        assert e1 <= tol and e2 <= tol
        if prec == "fp16x3":
            assert np.abs(probs - gold).max() <= 1e-3
            assert np.array_equal(probs.argmax(1), gold.argmax(1))
        plan.capture()
        assert np.array_equal(plan.forward(x).cpu().numpy(), probs)


def test_unet_probabilities(built_lib):
    import torch

    from oracle.torch_reference import unet_forward_reference, unet_infer_tail_reference
    from tiatoolbox_b200.models.resnet_plan import UNetPlan
    from tiatoolbox_b200.synthetic import he_tiles

    _no_tf32()
    sd = unet_test_weights()
    tiles = he_tiles(2, 256, 256, seed=0)
    x = torch.from_numpy(tiles).cuda()
    gold = np.load(GOLDEN / "unet_golden.npz")["probs"]
    with torch.inference_mode():
        ref = unet_infer_tail_reference(unet_forward_reference({k: v.cuda() for k, v in sd.items()}, x.permute(0, 3, 1, 2).float())).cpu().numpy()
    plan = UNetPlan(sd, batch=2, device="cuda:0", in_size=256, precision="fp16x3")
    probs = plan.forward(x).cpu().numpy()
    assert probs.shape == (2, 128, 128, 2)
    e1, e2 = rel_l2(probs, ref), rel_l2(probs, gold)
    print(f"unet fp16x3: rel_l2 vs torch fp32 {e1:.2e}, vs reference golden {e2:.2e}")
    assert e1 <= 1e-3 and e2 <= 1e-3
    print(f"unet max |dprob| vs reference golden {np.abs(probs - gold).max():.2e}")
    assert np.abs(probs - gold).max() <= 1e-3          # probabilities in [0, 1]: north_star's 1e-3
    assert (probs.argmax(-1) == gold.argmax(-1)).mean() >= 0.999


def test_unet_at_the_real_tile_size_1024(built_lib):
    """BASELINE config 3 (`fcn-tissue_mask`, pretrained_model.yaml:572-592: 1024 x 1024 in, 512 x 512 out): plan construction, memory and
    parity at the real tile size, through the public `UNetModel.infer_batch`, against the UNMODIFIED reference's CPU fp32 output on the same
    seeded tile (tests/golden/unet1024_golden.npz, oracle/make_golden_cnn.py::make_unet_1024; every 4th pixel)."""
    import torch

    from tiatoolbox_b200.models.architecture.unet import UNetModel
    from tiatoolbox_b200.synthetic import he_tiles

    gold = np.load(GOLDEN / "unet1024_golden.npz")
    model = UNetModel(3, 2, "resnet50", decoder_block=[3])
    model.load_state_dict(unet_test_weights())
    model.to("cuda:0")
    tile = he_tiles(1, 1024, 1024, seed=42)
    probs = UNetModel.infer_batch(model, torch.from_numpy(tile), device="cuda:0")
    assert probs.shape == (1, 512, 512, 2) and probs.dtype == np.float32
    e = rel_l2(probs[:, ::4, ::4], gold["probs_stride4"])
    m = float(np.abs(probs[:, ::4, ::4] - gold["probs_stride4"]).max())
    print(f"unet 1024: rel_l2 {e:.2e}  max |dprob| {m:.2e}  class-1 fraction {float((probs.argmax(-1) == 1).mean()):.4f} (reference {float(gold['frac_class1']):.4f})")
    assert e <= 1e-3 and m <= 1e-3
    assert np.allclose(probs.mean(axis=(0, 1, 2)), gold["mean"], atol=1e-4)
    assert abs(float((probs.argmax(-1) == 1).mean()) - float(gold["frac_class1"])) <= 1e-3
    # batch of two tiles through the engine contract: (N, 512, 512) bool predictions
    from tiatoolbox_b200.models.engine import SemanticSegmentor

    eng = SemanticSegmentor(model, batch_size=2, device="cuda:0")
    out = eng.run(np.concatenate([tile, tile[:, ::-1].copy()]))
    assert out["predictions"].shape == (2, 512, 512) and out["predictions"].dtype == bool
    assert np.array_equal(out["predictions"][0], probs[0].argmax(-1).astype(bool))
