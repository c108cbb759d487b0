"""WSI mode (BASELINE config 5 data flow) on the device: strided patch views of the uploaded slide band, the device stitcher writing
into the head canvases, batched on-device post-processing of the tile sets, and the ordered merge — against the reference's tile-mode
flow restated with the oracles (tests/test_wsi_engine_cpu.expected_result)."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import pytest

from tests.test_wsi_engine_cpu import IO_SMALL, check, decode_heads, expected_result, synthetic_slide

pytestmark = pytest.mark.gpu


def _decoder_engine(batch_size: int = 7):
    """A NucleusInstanceSegmentor whose network is the pixel-wise decoder of the synthetic slide (torch ops on the device); stitching and
    post-processing are the real CUDA kernels of libtiab200.so."""
    import torch

    from tiatoolbox_b200.models.engine.nucleus_instance_segmentor import NucleusInstanceSegmentor

    class DecoderEngine(NucleusInstanceSegmentor):
        def __init__(self) -> None:  # noqa: D107
            self.batch_size = batch_size
            self.device = torch.device("cuda", 0)
            self.ioconfig = IO_SMALL
            self.model = SimpleNamespace(_plan_batch=None)
            self.return_probabilities = False

        def _wsi_forward(self, x):
            assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
            c = x[:, 12:52, 12:52].long()
            # per-level tables computed by numpy (torch's CUDA division by a constant is a multiplication by its reciprocal, 1 ulp off)
            lv = np.arange(256, dtype=np.uint8)[:, None].repeat(3, 1)
            t_np, t_hv, _ = decode_heads(lv)
            lut_np = torch.from_numpy(np.ascontiguousarray(t_np[:, 0])).to(x.device)
            lut_hv = torch.from_numpy(np.ascontiguousarray(t_hv[:, 0])).to(x.device)
            npm = lut_np[c[..., 0:1]]
            hv = lut_hv[c[..., 1:3]]
            tp = torch.where(npm > 0.5, 1.0 + ((c[..., 1:2] // 16) % 5).float(), torch.zeros_like(npm))
            return [npm.contiguous(), hv.contiguous(), tp.contiguous()]

    return DecoderEngine()


def test_wsi_mode_device_stitch_and_postproc_match_the_oracle_flow(built_lib):
    image = synthetic_slide(390, 450, seed=3)
    kept, wsi_pred, heads = expected_result(image, IO_SMALL)
    eng = _decoder_engine()
    eng.return_probabilities = True
    out = eng.run(image, patch_mode=False, return_predictions=(True,))
    for got, want in zip(out["probabilities"], decode_heads(image)):
        assert np.array_equal(got, want)                    # device stitcher == pixel-wise decode of the slide, bit for bit
    check(out, kept, wsi_pred)                              # merged instances + WSI instance map == oracle flow, bit for bit


def test_wsi_mode_sliver_tiles_and_single_tile_region(built_lib):
    # 325 = 2 x 160 + 5: the last tile column is a 5-px sliver (below the post-processing kernels' minimum side)
    image = synthetic_slide(200, 325, seed=5, n_nuclei=50)
    out = _decoder_engine().run(image, patch_mode=False, return_predictions=True)
    assert out["predictions"].shape == (200, 325) and len(out["box"]) > 5
    small = image[:120, :150]
    from oracle import hovernet_postproc_oracle as ppo

    out = _decoder_engine().run(small, patch_mode=False, return_predictions=True)
    inst, info = ppo.postproc(*decode_heads(small))
    assert np.array_equal(out["predictions"], inst)
    assert len(out["box"]) == len(info) and all(np.array_equal(a, r["box"]) for a, r in zip(out["box"], info.values()))


def test_wsi_mode_hovernet_vahadane_end_to_end(built_lib):
    """Config 5 in small: Vahadane-normalised patches -> hovernet_fast -> WSI head maps -> tile-mode post-processing.  The WSI head maps
    must equal the per-patch `infer_batch` outputs placed on the output grid, and the instance result must equal the oracle flow
    applied to those head maps."""
    import torch

    from oracle import hovernet_postproc_oracle as ppo
    from oracle import wsi_oracle as wo
    from tests.cnn_common import hovernet_test_weights
    from tiatoolbox_b200.models.architecture import PRETRAINED_INFO
    from tiatoolbox_b200.models.engine.nucleus_instance_segmentor import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_target, he_tiles
    from tiatoolbox_b200.tools.patchextraction import PatchExtractor
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    io = PRETRAINED_INFO["hovernet_fast-pannuke"]["ioconfig"]
    H, W = 1100, 1250
    tiles = he_tiles(25, 256, 256, seed=11)
    image = np.concatenate([np.concatenate(list(tiles[r * 5:(r + 1) * 5]), axis=1) for r in range(5)], axis=0)[:H, :W].copy()
    norm = get_normalizer("vahadane")
    norm.fit(he_target(256, 256))
    eng = NucleusInstanceSegmentor("hovernet_fast-pannuke", batch_size=8, device="cuda")
    eng.model.load_state_dict(hovernet_test_weights("fast", 6))
    eng.model.preproc_func = norm.transform
    out = eng.run(image, patch_mode=False, ioconfig=io, return_predictions=True, return_probabilities=True)
    torch.cuda.synchronize()
    # ---- head canvases == per-patch outputs on the grid
    ins, outs = PatchExtractor.get_coordinates(image_shape=(W, H), patch_input_shape=io.patch_input_shape,
                                               patch_output_shape=io.patch_output_shape, stride_shape=io.stride_shape)
    assert np.array_equal(out["coordinates"], ins)
    padded = np.full((H + 512, W + 512, 3), 255, np.uint8)
    padded[256:256 + H, 256:256 + W] = image
    patches = np.stack([padded[y0 + 256:y1 + 256, x0 + 256:x1 + 256] for x0, y0, x1, y1 in ins])
    per_patch = eng.run(patches, patch_mode=True, return_probabilities=True)["probabilities"]
    for head, canvas in zip(per_patch, out["probabilities"]):
        want = np.zeros_like(canvas)
        for k, (x0, y0, x1, y1) in enumerate(outs):
            yy, xx = min(y1, H), min(x1, W)
            want[y0:yy, x0:xx] = head[k][: yy - y0, : xx - x0]
        assert canvas.shape == (H, W, head.shape[-1])
        np.testing.assert_allclose(canvas, want, rtol=0, atol=1e-5)
    # ---- instances == oracle tile-mode flow on those canvases
    heads = out["probabilities"]
    tiles_, wsi_pred, mx = [], np.zeros((H, W), np.int32), None
    for bounds, flag, mode in wo.build_tile_tasks(wo.get_tile_info((W, H), io.tile_shape, io.patch_output_shape, io.margin)):
        x0, y0, x1, y1 = (int(v) for v in bounds)
        inst, info = ppo.postproc(*[hd[max(y0, 0):y1, max(x0, 0):x1] for hd in heads])
        b = np.array([x0, y0, x1, y1])
        xe, ye = min(x1, W), min(y1, H)
        prev = wsi_pred[y0:ye, x0:xe]
        new = inst[: ye - y0, : xe - x0]
        clash = np.setdiff1d(np.unique(new[(new > 0) & (prev > 0)]), [0])
        keep = (new > 0) & ~np.isin(new, clash)
        base = 0 if mx is None else mx
        prev[keep] = new[keep] + base
        mx = int(prev.max()) + base if keep.any() else base
        tiles_.append((b, flag, mode, list(info.values())))
    kept = wo.replay_merge(tiles_, io.margin)
    assert len(out.get("box", [])) == len(kept)
    for i, rec in enumerate(kept):
        assert np.array_equal(out["box"][i], rec["box"]) and np.array_equal(out["contours"][i], rec["contours"])
        assert out["type"][i] == rec["type"]
    assert np.array_equal(out["predictions"], wsi_pred)


def test_patch_predictor_wsi_mode(built_lib):
    """PatchPredictor on an in-memory region: the coordinate grid and 255 padding of the reference's WSI dataset, results == patch mode."""
    from tests.cnn_common import resnet18_test_weights
    from tiatoolbox_b200.models.engine import PatchPredictor
    from tiatoolbox_b200.synthetic import he_tiles
    from tiatoolbox_b200.tools.patchextraction import PatchExtractor

    eng = PatchPredictor("resnet18-kather100k", batch_size=4, device="cuda")
    eng.model.load_state_dict(resnet18_test_weights())
    t = he_tiles(6, 256, 256, seed=13)
    region = np.concatenate([np.concatenate(list(t[:3]), axis=1), np.concatenate(list(t[3:]), axis=1)], axis=0)[:500, :700].copy()
    out = eng.run(region, patch_mode=False, return_probabilities=True)
    coords = PatchExtractor.get_coordinates(image_shape=(700, 500), patch_input_shape=(224, 224), stride_shape=(224, 224))
    assert np.array_equal(out["coordinates"], coords) and len(coords) == 4 * 3
    padded = np.full((672, 896, 3), 255, np.uint8)
    padded[:500, :700] = region
    patches = np.stack([padded[y0:y1, x0:x1] for x0, y0, x1, y1 in coords])
    ref = eng.run(patches, return_probabilities=True)
    assert np.array_equal(out["probabilities"], ref["probabilities"]) and np.array_equal(out["predictions"], ref["predictions"])
