"""GPU tests of the engine API (run / predict, output dict contract of SURVEY.md §8b) on seeded weights."""
from __future__ import annotations

import numpy as np
import pytest

from tests.cnn_common import GOLDEN, hovernet_test_weights, resnet18_test_weights, unet_test_weights

pytestmark = pytest.mark.gpu


def test_patch_predictor_contract(built_lib):
    from tiatoolbox_b200.models.engine import PatchPredictor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = PatchPredictor("resnet18-kather100k", batch_size=8, device="cuda")
    eng.model.load_state_dict(resnet18_test_weights())
    tiles = he_tiles(16, 224, 224, seed=0)
    gold = np.load(GOLDEN / "resnet18_golden.npz")["probs"]
    out = eng.run(tiles[:13], patch_mode=True, return_probabilities=True, labels=list(range(13)))  # ragged last batch
    assert set(out) == {"predictions", "probabilities", "labels"}
    assert out["predictions"].shape == (13,) and out["probabilities"].shape == (13, 9)
    assert np.array_equal(out["predictions"], gold[:13].argmax(1))
    assert np.abs(out["probabilities"] - gold[:13]).max() <= 1e-3
    assert out["predictions"].dtype == np.uint8 or out["predictions"].dtype == bool
    out2 = eng.predict(list(tiles[:4]), return_probabilities=False)  # 1.x alias, list input
    assert "probabilities" not in out2 and np.array_equal(out2["predictions"], out["predictions"][:4])
    with pytest.raises(NotImplementedError):
        eng.run(tiles[:2], patch_mode=False)
    with pytest.raises(NotImplementedError):
        eng.run(tiles[:2], output_type="zarr")


def test_semantic_segmentor_contract(built_lib):
    from tiatoolbox_b200.models.engine import SemanticSegmentor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = SemanticSegmentor("fcn-tissue_mask", batch_size=2, device="cuda")
    eng.model.load_state_dict(unet_test_weights())
    tiles = he_tiles(2, 256, 256, seed=0)
    gold = np.load(GOLDEN / "unet_golden.npz")["probs"]
    out = eng.run(tiles, return_probabilities=True)
    assert out["predictions"].shape == (2, 128, 128) and out["predictions"].dtype == bool
    assert out["probabilities"].shape == (2, 128, 128, 2)
    assert (out["predictions"] == gold.argmax(-1).astype(bool)).mean() >= 0.999


def test_nucleus_instance_segmentor_matches_oracle_postproc(built_lib):
    from oracle import hovernet_postproc_oracle as po
    from tiatoolbox_b200.models.architecture.hovernet import HoVerNet
    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = NucleusInstanceSegmentor("hovernet_fast-pannuke", batch_size=2, device="cuda")
    eng.model.load_state_dict(hovernet_test_weights("fast", 6))
    tiles = he_tiles(3, 256, 256, seed=0)
    out = eng.run(tiles, return_probabilities=True)
    assert out["predictions"].shape == (3, 164, 164) and out["predictions"].dtype == np.int32
    for k in ("box", "centroid", "contours", "prob", "type"):
        assert len(out[k]) == 3
    np_m, hv_m, tp_m = out["probabilities"]
    assert np_m.shape == (3, 164, 164, 1) and hv_m.shape == (3, 164, 164, 2) and tp_m.shape == (3, 164, 164, 1)
    # same head maps through the reference-style static infer_batch
    heads = HoVerNet.infer_batch(eng.model, tiles[:2], device="cuda")
    assert np.array_equal(heads[0], np_m[:2])
    # the engine's instance maps / tables equal the reference post-processing of those head maps
    for i in range(3):
        ref_inst, ref_info = po.postproc(np_m[i], hv_m[i], tp_m[i])
        assert np.array_equal(out["predictions"][i], ref_inst)
        assert len(out["box"][i]) == len(ref_info)
        for j, rec in enumerate(ref_info.values()):
            assert np.array_equal(out["box"][i][j], rec["box"]) and out["type"][i][j] == rec["type"]
            assert np.array_equal(out["contours"][i][j], rec["contours"])
    # model.postproc on one image (reference API): same result
    res = eng.model.postproc([np_m[0], hv_m[0], tp_m[0]])[0]
    assert res["task_type"] == "nuclei_segmentation" and res["seg_type"] == "instance"
    assert np.array_equal(res["predictions"], out["predictions"][0])


def test_stain_preproc_func_runs_on_device(built_lib):
    from tiatoolbox_b200.models.engine import PatchPredictor
    from tiatoolbox_b200.synthetic import he_target, he_tiles
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    norm = get_normalizer("macenko")
    norm.fit(he_target(256, 256))
    eng = PatchPredictor("resnet18-kather100k", batch_size=4, device="cuda")
    eng.model.load_state_dict(resnet18_test_weights())
    tiles = he_tiles(6, 224, 224, seed=2)
    eng.model.preproc_func = norm.transform           # the slide-graph notebook wiring
    a = eng.run(tiles, return_probabilities=True)["probabilities"]
    eng.model.preproc_func = None
    b = eng.run(norm.transform_batch(tiles), return_probabilities=True)["probabilities"]
    assert np.array_equal(a, b)
    c = eng.run(tiles, return_probabilities=True)["probabilities"]
    assert np.abs(a - c).max() > 1e-4  # normalisation did change the input
