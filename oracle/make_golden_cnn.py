"""CNN fixtures: (i) calibrated BatchNorm statistics for the seeded test weights, (ii) outputs of the
UNMODIFIED reference modules (loaded through oracle/ref_shim.py) for those weights on seeded tiles.

TEST INFRASTRUCTURE ONLY; run here (needs /root/reference):  python -m oracle.make_golden cnn
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _f16(a: torch.Tensor) -> np.ndarray:
    return a.detach().cpu().numpy().astype(np.float32)


def make_hovernet() -> None:
    from oracle import ref_shim
    from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
    from oracle.torch_reference import hovernet_forward_reference
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec
    from tiatoolbox_b200.synthetic import he_tiles

    m = ref_shim.load_models()
    out = {}
    for tag, mode, nt, size in (("fast", "fast", 6, 256),):
        sd = seeded_state_dict(hovernet_spec(3, nt, mode), seed=0)
        xcal = torch.from_numpy(he_tiles(2, size, size, seed=100)).permute(0, 3, 1, 2).float()
        stats = calibrate(sd, lambda s, x: hovernet_forward_reference(s, x, mode, nt), xcal)
        np.savez_compressed(GOLDEN / f"hovernet_{tag}_bn_calib.npz", **{k: v.numpy() for k, v in stats.items()})
        net = m.hovernet.HoVerNet(num_types=nt, mode=mode).eval()
        net.load_state_dict(sd)
        tiles = he_tiles(2, size, size, seed=0)
        with torch.inference_mode():
            raw = net(torch.from_numpy(tiles).permute(0, 3, 1, 2).float())
        heads = m.hovernet.HoVerNet.infer_batch(net, torch.from_numpy(tiles), device="cpu")
        for k, v in raw.items():  # raw branch outputs, every 2nd pixel (keeps the fixture ~1 MB)
            out[f"{tag}_logits_{k}"] = _f16(v.permute(0, 2, 3, 1))[:, ::2, ::2]
        for k, v in zip(("np", "hv", "tp"), heads):  # infer_batch heads of tile 0, full resolution
            out[f"{tag}_head_{k}"] = np.asarray(v, dtype=np.float32)[:1]
    np.savez_compressed(GOLDEN / "hovernet_golden.npz", **out)
    print("hovernet golden:", {k: v.shape for k, v in out.items()})


def synthetic_layer_map(h: int = 400, w: int = 420, seed: int = 0) -> np.ndarray:
    """A 5-class layer map (0 background, 1 other tissue, 2 basal, 3 core epithelium, 4 keratin) made of bands + islands + speckle, as the
    float32 argmax head HoVerNetPlus.infer_batch returns (H, W, 1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), np.uint8)
    m[40:, :] = 1
    wave = (30 * np.sin(xx / 40.0)).astype(int)
    m[(yy > 120 + wave) & (yy < 330 + wave)] = 3
    m[(yy > 120 + wave) & (yy < 150 + wave)] = 2
    m[(yy > 290 + wave) & (yy < 330 + wave)] = 4
    m[60:80, 30:60] = 3            # small epithelium island (< 20 000 px): removed
    m[200:215, 100:118] = 1        # hole in the epithelium smaller than the 20x20 closing
    speck = rng.random((h, w)) < 0.01
    m[speck] = rng.integers(0, 5, int(speck.sum()))
    return m.astype(np.float32)[..., None]


def make_hovernetplus() -> None:
    """HoVerNetPlus (hovernetplus-oed shape: 3 nucleus types, 5 layers): BN calibration, reference logits / heads, reference postproc."""
    import importlib
    import sys

    from oracle import hovernet_postproc_oracle as po
    from oracle import ref_shim
    from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
    from oracle.torch_reference import hovernet_forward_reference
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec
    from tiatoolbox_b200.synthetic import he_tiles, hovernet_head_maps

    ref_shim.load_models()
    hp = importlib.import_module("tiatoolbox.models.architecture.hovernetplus")
    nt, nl = 3, 5
    sd = seeded_state_dict(hovernet_spec(3, nt, "fast", nl), seed=0)
    xcal = torch.from_numpy(he_tiles(2, 256, 256, seed=100)).permute(0, 3, 1, 2).float()
    stats = calibrate(sd, lambda s, x: hovernet_forward_reference(s, x, "fast", nt, num_layers=nl), xcal)
    np.savez_compressed(GOLDEN / "hovernetplus_bn_calib.npz", **{k: v.numpy() for k, v in stats.items()})
    net = hp.HoVerNetPlus(num_types=nt, num_layers=nl).eval()
    net.load_state_dict(sd)
    tiles = he_tiles(2, 256, 256, seed=0)
    out = {}
    with torch.inference_mode():
        raw = net(torch.from_numpy(tiles).permute(0, 3, 1, 2).float())
    heads = hp.HoVerNetPlus.infer_batch(net, torch.from_numpy(tiles), device="cpu")
    for k, v in raw.items():
        out[f"logits_{k}"] = _f16(v.permute(0, 2, 3, 1))[:, ::2, ::2]
    for k, v in zip(("np", "hv", "tp", "ls"), heads):
        out[f"head_{k}"] = np.asarray(v, dtype=np.float32)[:1]
    # reference post-processing (two restated scikit-image functions injected) on synthetic head maps + a synthetic layer map
    hn = ref_shim.load_models().hovernet
    hn.remove_small_objects, hn.watershed = po.remove_small_objects, po.watershed
    sys.modules["skimage.morphology"].remove_small_objects = po.remove_small_objects_bool
    npm, hvm, tpm = hovernet_head_maps(1, 164, seed=4)
    ls = synthetic_layer_map()
    out["pp_ls_in"] = ls
    out["pp_layer"] = hp.HoVerNetPlus._proc_ls(ls)
    out["pp_inst"] = hp.HoVerNetPlus._proc_np_hv(npm[0], hvm[0], scale_factor=0.5)
    linfo = hp.HoVerNetPlus._get_layer_info(out["pp_layer"], (3, 7))
    out["pp_layer_types"] = np.array([int(r["type"]) for r in linfo.values()])
    out["pp_layer_boxes"] = np.array([r["box"] for r in linfo.values()])
    out["pp_layer_contour_lens"] = np.array([len(r["contours"]) for r in linfo.values()])
    out["pp_layer_contour0"] = np.asarray(next(iter(linfo.values()))["contours"])
    np.savez_compressed(GOLDEN / "hovernetplus_golden.npz", **out)
    print("hovernetplus golden:", {k: v.shape for k, v in out.items()}, "layer classes", np.unique(out["pp_layer"]), "instances",
          int(out["pp_inst"].max()))


def make_cnn() -> None:
    GOLDEN.mkdir(parents=True, exist_ok=True)
    make_hovernet()
    make_resnet18()
    make_unet()


def make_postproc() -> None:
    pass


def make_resnet18() -> None:
    from oracle import ref_shim
    from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
    from oracle.torch_reference import cnn_forward_reference
    from tiatoolbox_b200.models.architecture.vanilla import cnn_spec
    from tiatoolbox_b200.synthetic import he_tiles

    m = ref_shim.load_models()
    sd = seeded_state_dict(cnn_spec("resnet18", 9), seed=0)
    xcal = torch.from_numpy(he_tiles(8, 224, 224, seed=100)).permute(0, 3, 1, 2).float() / 255.0
    stats = calibrate(sd, lambda s, x: cnn_forward_reference(s, x, "resnet18"), xcal)
    np.savez_compressed(GOLDEN / "resnet18_bn_calib.npz", **{k: v.numpy() for k, v in stats.items()})
    net = m.vanilla.CNNModel("resnet18", 9).eval()
    net.load_state_dict(sd)
    tiles = he_tiles(16, 224, 224, seed=0)
    # the engine feeds ToTensor-scaled float32 NHWC in [0, 1] (dataset/classification.py:30-63)
    probs = m.vanilla.CNNModel.infer_batch(net, torch.from_numpy(tiles.astype(np.float32) / 255.0), device="cpu")
    np.savez_compressed(GOLDEN / "resnet18_golden.npz", probs=np.asarray(probs, np.float32))
    print("resnet18 golden:", probs.shape, probs[0])


def make_unet() -> None:
    from oracle import ref_shim
    from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
    from oracle.torch_reference import unet_forward_reference
    from tiatoolbox_b200.models.architecture.unet import unet_spec
    from tiatoolbox_b200.synthetic import he_tiles

    m = ref_shim.load_models()
    sd = seeded_state_dict(unet_spec(3, 2, (3,)), seed=0)
    xcal = torch.from_numpy(he_tiles(2, 256, 256, seed=100)).permute(0, 3, 1, 2).float()
    stats = calibrate(sd, lambda s, x: unet_forward_reference(s, x), xcal)
    np.savez_compressed(GOLDEN / "unet_bn_calib.npz", **{k: v.numpy() for k, v in stats.items()})
    net = m.unet.UNetModel(3, 2, "resnet50", decoder_block=[3]).eval()
    net.load_state_dict(sd)
    tiles = he_tiles(2, 256, 256, seed=0)
    probs = m.unet.UNetModel.infer_batch(net, torch.from_numpy(tiles), device="cpu")
    np.savez_compressed(GOLDEN / "unet_golden.npz", probs=np.asarray(probs, np.float32))
    print("unet golden:", probs.shape, probs.mean(axis=(0, 1, 2)))


def make_unet_1024() -> None:
    """BASELINE config 3 at its real tile size: the UNMODIFIED reference UNetModel.infer_batch on one seeded 1024 x 1024 tile (CPU fp32),
    every 4th output pixel (512 x 512 x 2 -> 128 x 128 x 2 keeps the fixture at ~130 KB).   python -c "from oracle.make_golden_cnn import make_unet_1024 as f; f()" """
    from oracle import ref_shim
    from tests.cnn_common import unet_test_weights
    from tiatoolbox_b200.synthetic import he_tiles

    m = ref_shim.load_models()
    net = m.unet.UNetModel(3, 2, "resnet50", decoder_block=[3]).eval()
    net.load_state_dict(unet_test_weights())
    tile = he_tiles(1, 1024, 1024, seed=42)
    probs = np.asarray(m.unet.UNetModel.infer_batch(net, torch.from_numpy(tile), device="cpu"), np.float32)
    assert probs.shape == (1, 512, 512, 2)
    np.savez_compressed(GOLDEN / "unet1024_golden.npz", probs_stride4=probs[:, ::4, ::4], mean=probs.mean(axis=(0, 1, 2)),
                        frac_class1=np.float64((probs.argmax(-1) == 1).mean()))
    print("unet 1024 golden:", probs.shape, probs.mean(axis=(0, 1, 2)), (probs.argmax(-1) == 1).mean())


def make_hovernet_original() -> None:
    """hovernet `mode="original"` (270 x 270 in, 80 x 80 out, 5 x 5 decoder convolutions): BN calibration on the CPU + the UNMODIFIED reference
    module's branch logits and `infer_batch` heads on one seeded tile.   python -c "from oracle.make_golden_cnn import make_hovernet_original as f; f()" """
    from oracle import ref_shim
    from oracle.torch_reference import hovernet_forward_reference
    from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec
    from tiatoolbox_b200.synthetic import he_tiles

    m = ref_shim.load_models()
    sd = seeded_state_dict(hovernet_spec(3, 6, "original"), seed=0)
    xcal = torch.from_numpy(he_tiles(2, 270, 270, seed=100)).permute(0, 3, 1, 2).float()
    stats = calibrate(sd, lambda s, x: hovernet_forward_reference(s, x, "original", 6), xcal)
    np.savez_compressed(GOLDEN / "hovernet_original_bn_calib.npz", **{k: v.numpy() for k, v in stats.items()})
    sd.update(stats)
    net = m.hovernet.HoVerNet(num_types=6, mode="original").eval()
    net.load_state_dict(sd)
    tile = he_tiles(1, 270, 270, seed=2)
    with torch.inference_mode():
        raw = net(torch.from_numpy(tile).permute(0, 3, 1, 2).float())
    heads = m.hovernet.HoVerNet.infer_batch(net, torch.from_numpy(tile), device="cpu")
    out = {f"logits_{k}": _f16(v.permute(0, 2, 3, 1)) for k, v in raw.items()}
    out.update({f"head_{k}": np.asarray(v, np.float32) for k, v in zip(("np", "hv", "tp"), heads)})
    np.savez_compressed(GOLDEN / "hovernet_original_golden.npz", **out)
    print("hovernet original golden:", {k: v.shape for k, v in out.items()})
