"""HoVerNetPlus without a GPU: the fp32 restatement of the forward against the outputs of the unmodified reference module, the host-side
layer post-processing (`_proc_ls`, `_get_layer_info`) and the scale-0.5 nucleus oracle against the reference's outputs
(tests/golden/hovernetplus_golden.npz, oracle/make_golden_cnn.make_hovernetplus), state-dict key parity, registry entry."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import hovernet_postproc_oracle as po
from oracle import ref_shim
from tests.cnn_common import GOLDEN, rel_l2

GOLD = np.load(GOLDEN / "hovernetplus_golden.npz")


def hovernetplus_test_weights():
    from tiatoolbox_b200.models.architecture._params import seeded_state_dict, with_calibration
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec

    return with_calibration(seeded_state_dict(hovernet_spec(3, 3, "fast", 5), seed=0), GOLDEN / "hovernetplus_bn_calib.npz")


def test_reference_forward_restatement_matches_reference_module():
    from oracle.torch_reference import hovernet_forward_reference, hovernet_heads_reference
    from tiatoolbox_b200.synthetic import he_tiles

    sd = hovernetplus_test_weights()
    x = torch.from_numpy(he_tiles(2, 256, 256, seed=0)[:1]).permute(0, 3, 1, 2).float()
    with torch.inference_mode():
        out = hovernet_forward_reference(sd, x, "fast", 3, num_layers=5)
    assert list(out) == ["tp", "np", "hv", "ls"]
    for k, v in out.items():
        got = v.permute(0, 2, 3, 1).numpy()[:, ::2, ::2]
        assert got.shape[-1] == {"tp": 3, "np": 2, "hv": 2, "ls": 5}[k]
        assert rel_l2(got, GOLD[f"logits_{k}"][:1]) < 1e-5, k
    heads = hovernet_heads_reference(out)
    assert len(heads) == 4
    for name, h in zip(("np", "hv", "tp", "ls"), heads):
        g = GOLD[f"head_{name}"]
        if name in ("tp", "ls"):
            assert (h.numpy() == g).mean() > 0.999, name
        else:
            assert rel_l2(h.numpy(), g) < 1e-5, name


def test_layer_postprocessing_matches_reference_outputs():
    from tiatoolbox_b200.models.architecture.hovernetplus import HoVerNetPlus

    layer = HoVerNetPlus._proc_ls(GOLD["pp_ls_in"])
    assert layer.dtype == np.uint8 and np.array_equal(layer, GOLD["pp_layer"])
    assert np.array_equal(po.proc_ls(GOLD["pp_ls_in"]), GOLD["pp_layer"])
    for impl in (HoVerNetPlus._get_layer_info, po.get_layer_info):
        info = impl(layer, (3, 7))
        assert [int(r["type"]) for r in info.values()] == GOLD["pp_layer_types"].tolist()
        assert np.array_equal(np.array([r["box"] for r in info.values()]), GOLD["pp_layer_boxes"])
        assert [len(r["contours"]) for r in info.values()] == GOLD["pp_layer_contour_lens"].tolist()
        assert np.array_equal(next(iter(info.values()))["contours"], GOLD["pp_layer_contour0"])
    assert list(info.keys()) == list(range(1, len(info) + 1))
    assert HoVerNetPlus._get_layer_info(np.zeros((30, 30), np.uint8)) == {}


def test_scale_half_nucleus_oracle_matches_reference_output():
    from tiatoolbox_b200.synthetic import hovernet_head_maps

    npm, hvm, _ = hovernet_head_maps(1, 164, seed=4)
    inst = po.proc_np_hv(npm[0], hvm[0], scale_factor=0.5)
    assert np.array_equal(inst, GOLD["pp_inst"]) and inst.max() > 10
    assert not np.array_equal(inst, po.proc_np_hv(npm[0], hvm[0]))      # the scale does change the result


def test_registry_and_constructor():
    from tiatoolbox_b200.models.architecture import get_pretrained_model
    from tiatoolbox_b200.models.architecture.hovernetplus import HoVerNetPlus

    model, io = get_pretrained_model("hovernetplus-oed")
    assert isinstance(model, HoVerNetPlus) and model.tasks == ["nuclei_segmentation", "layer_segmentation"]
    assert model.num_types == 3 and model.num_layers == 5 and model.mode == "fast"
    assert model.class_dict["layer_segmentation"][4] == "Keratin" and tuple(io.tile_shape) == (2048, 2048) and io.margin == 128
    with pytest.raises(ValueError, match="num_types and num_layers"):
        HoVerNetPlus(num_types=3)


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present (GPU box)")
def test_state_dict_keys_equal_the_reference_module():
    import importlib

    from tiatoolbox_b200.models.architecture.hovernetplus import HoVerNetPlus

    ref_shim.load_models()
    hp = importlib.import_module("tiatoolbox.models.architecture.hovernetplus")
    ref = {k: tuple(v.shape) for k, v in hp.HoVerNetPlus(num_types=3, num_layers=5).state_dict().items()}
    mine = {k: tuple(v.shape) for k, v in HoVerNetPlus(num_types=3, num_layers=5)._sd.items()}
    assert ref == mine and len(ref) == 964
