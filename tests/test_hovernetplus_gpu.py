"""HoVerNetPlus on the device: the four-branch plan against the outputs of the unmodified reference module, the scale-0.5 nucleus
post-processing (11-tap Sobel, 3-pixel marker threshold) bit-exactly against the oracle, and the two-task engine contract."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import hovernet_postproc_oracle as po
from tests.cnn_common import GOLDEN, rel_l2
from tests.test_hovernetplus_cpu import hovernetplus_test_weights

pytestmark = pytest.mark.gpu
GOLD = np.load(GOLDEN / "hovernetplus_golden.npz")


def test_postproc_scale_half_bit_exact(built_lib):
    import torch

    from tiatoolbox_b200.models import postproc as pp
    from tiatoolbox_b200.synthetic import hovernet_head_maps

    npm, hvm, tpm = hovernet_head_maps(3, 164, seed=4)
    inst, table = pp.proc_np_hv_batch(torch.from_numpy(npm).cuda(), torch.from_numpy(hvm).cuda(), torch.from_numpy(tpm).cuda(), scale_factor=0.5)
    inst = inst.cpu().numpy()
    assert np.array_equal(inst[0], GOLD["pp_inst"])                      # == the reference's own _proc_np_hv(scale_factor=0.5)
    for i in range(3):
        assert np.array_equal(inst[i], po.proc_np_hv(npm[i], hvm[i], scale_factor=0.5)), i
    inst1, _ = pp.proc_np_hv_batch(torch.from_numpy(npm).cuda(), torch.from_numpy(hvm).cuda(), None)
    for i in range(3):                                                     # scale 1 is untouched by the generalisation
        assert np.array_equal(inst1[i].cpu().numpy(), po.proc_np_hv(npm[i], hvm[i])), i
    with pytest.raises(Exception, match="scale_factor"):
        pp.proc_np_hv_batch(torch.from_numpy(npm).cuda(), torch.from_numpy(hvm).cuda(), None, scale_factor=0.25)


def test_plan_forward_matches_reference_module(built_lib):
    import torch

    from tiatoolbox_b200.models.architecture.hovernetplus import HoVerNetPlus
    from tiatoolbox_b200.synthetic import he_tiles

    model = HoVerNetPlus(num_types=3, num_layers=5)
    model.load_state_dict(hovernetplus_test_weights())
    model.to("cuda")
    tiles = he_tiles(2, 256, 256, seed=0)
    heads = HoVerNetPlus.infer_batch(model, tiles, device="cuda")
    assert len(heads) == 4 and heads[3].shape == (2, 164, 164, 1) and heads[3].dtype == np.float32
    plan = model.plan_for(2, torch.device("cuda", 0), (256, 256))
    for k in ("tp", "np", "hv", "ls"):
        got = plan.logits[k].cpu().numpy()[:, ::2, ::2]
        assert rel_l2(got, GOLD[f"logits_{k}"]) < 1e-3, (k, rel_l2(got, GOLD[f"logits_{k}"]))
    assert rel_l2(heads[0][:1], GOLD["head_np"]) < 1e-3 and rel_l2(heads[1][:1], GOLD["head_hv"]) < 1e-3
    assert (heads[2][:1] == GOLD["head_tp"]).mean() > 0.995 and (heads[3][:1] == GOLD["head_ls"]).mean() > 0.995


def test_two_task_engine_contract(built_lib):
    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = NucleusInstanceSegmentor("hovernetplus-oed", batch_size=2, device="cuda")
    eng.model.load_state_dict(hovernetplus_test_weights())
    tiles = he_tiles(3, 256, 256, seed=0)
    out = eng.run(tiles, return_probabilities=True)
    assert set(out) == {"nuclei_segmentation", "layer_segmentation", "probabilities"}
    nuc, lay = out["nuclei_segmentation"], out["layer_segmentation"]
    assert nuc["predictions"].shape == (3, 164, 164) and nuc["predictions"].dtype == np.int32
    assert lay["predictions"].shape == (3, 164, 164) and lay["predictions"].dtype == np.uint8
    assert all(len(lay[k]) == 3 for k in ("box", "contours", "type")) and all(len(nuc[k]) == 3 for k in ("box", "centroid", "contours", "prob", "type"))
    np_m, hv_m, tp_m, ls_m = out["probabilities"]
    for i in range(3):   # both tasks equal the oracle post-processing of the engine's own head maps
        inst, info, layer, linfo = po.postproc_plus(np_m[i], hv_m[i], tp_m[i], ls_m[i])
        assert np.array_equal(nuc["predictions"][i], inst) and len(nuc["box"][i]) == len(info)
        assert np.array_equal(lay["predictions"][i], layer) and len(lay["type"][i]) == len(linfo)
    res = eng.model.postproc([np_m[0], hv_m[0], tp_m[0], ls_m[0]])
    assert [r["task_type"] for r in res] == ["nuclei_segmentation", "layer_segmentation"] and res[1]["seg_type"] == "semantic"
    assert np.array_equal(res[0]["predictions"], nuc["predictions"][0]) and np.array_equal(res[1]["predictions"], lay["predictions"][0])


def test_two_task_wsi_mode_region(built_lib):
    """HoVerNetPlus on an in-memory region (fits one 2048-px tile: the un-tiled branch): nested two-task output == the oracle post-processing
    of the stitched head maps."""
    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = NucleusInstanceSegmentor("hovernetplus-oed", batch_size=4, device="cuda")
    eng.model.load_state_dict(hovernetplus_test_weights())
    t = he_tiles(6, 256, 256, seed=3)
    region = np.concatenate([np.concatenate(list(t[:3]), axis=1), np.concatenate(list(t[3:]), axis=1)], axis=0)[:420, :600].copy()
    out = eng.run(region, patch_mode=False, return_predictions=True, return_probabilities=True)
    assert set(out) >= {"nuclei_segmentation", "layer_segmentation", "coordinates", "probabilities"}
    heads = out["probabilities"]
    assert len(heads) == 4 and heads[3].shape == (420, 600, 1)
    inst, info, layer, linfo = po.postproc_plus(*heads)
    nuc, lay = out["nuclei_segmentation"], out["layer_segmentation"]
    assert np.array_equal(nuc["predictions"], inst) and len(nuc.get("box", [])) == len(info)
    assert np.array_equal(lay["predictions"], layer) and len(lay.get("box", [])) == len(linfo)
