"""GPU parity of the on-device HoVer-Net post-processing against the oracle
(oracle/hovernet_postproc_oracle.py = reference code path with the two scikit-image functions restated).

Integer / index outputs must be bit-exact: the foreground mask, the seed markers, the instance map
and the instance table's ids / boxes / types.  float64 intermediates (Sobel, distance map) follow
OpenCV's summation order and are compared bit-exactly too on this OpenCV build; centroids / type
probabilities to 1e-12."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import hovernet_postproc_oracle as po
from tiatoolbox_b200.synthetic import hovernet_head_maps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp(built_lib):
    from tiatoolbox_b200.models import postproc

    return postproc


def _run(pp, npm, hvm, tpm):
    import torch

    inst, table = pp.proc_np_hv_batch(torch.from_numpy(npm).cuda(), torch.from_numpy(hvm).cuda(),
                                      None if tpm is None else torch.from_numpy(tpm).cuda())
    torch.cuda.synchronize()
    return inst.cpu().numpy(), {k: v.cpu().numpy() for k, v in table.items()}


def test_instance_maps_and_intermediates_bit_exact(pp):
    npm, hvm, tpm = hovernet_head_maps(6, 164, seed=0)
    inst, table = _run(pp, npm, hvm, tpm)
    n = len(npm)
    blb = pp.fetch_debug(0, (n, 164, 164), np.uint8)
    hdir = pp.fetch_debug(1, (n, 164, 164), np.float32)
    sobh = pp.fetch_debug(3, (n, 164, 164), np.float64)
    dist = pp.fetch_debug(5, (n, 164, 164), np.float64)
    marker = pp.fetch_debug(6, (n, 164, 164), np.int32)
    import cv2

    for i in range(n):
        dbg = {}
        ref = po.proc_np_hv(npm[i], hvm[i], debug=dbg)
        assert np.array_equal(blb[i], dbg["blb"].astype(np.uint8)), i
        assert np.array_equal(hdir[i], dbg["h_dir"]), i
        assert np.array_equal(sobh[i], cv2.Sobel(dbg["h_dir"], cv2.CV_64F, 1, 0, ksize=21)), i
        assert np.array_equal(dist[i], dbg["dist"]), i
        assert np.array_equal(marker[i], dbg["marker"]), i
        assert np.array_equal(inst[i], ref), (i, (inst[i] != ref).sum())
        assert inst[i].max() > 10  # "must have some nuclei"


def test_instance_table_matches_get_instance_info(pp):
    npm, hvm, tpm = hovernet_head_maps(3, 164, seed=5)
    inst, table = _run(pp, npm, hvm, tpm)
    for i in range(len(npm)):
        ref_inst, ref_info = po.postproc(npm[i], hvm[i], tpm[i], offset=(7, 3))
        row = {k: v[i] for k, v in table.items()}
        info = pp.instance_info(inst[i], row, offset=(7, 3))
        assert list(info.keys()) == list(ref_info.keys())
        for k, r in ref_info.items():
            g = info[k]
            assert np.array_equal(g["box"], r["box"])
            assert np.allclose(g["centroid"], r["centroid"], rtol=0, atol=1e-12)
            assert np.array_equal(g["contours"], r["contours"])
            assert g["type"] == r["type"]
            assert abs(g["prob"] - r["prob"]) <= 1e-12


def test_edge_cases(pp):
    rng = np.random.default_rng(0)
    # no nuclei at all / everything foreground / tiny specks only / no type head
    npm = np.zeros((4, 64, 80, 1), np.float32)
    hvm = rng.normal(0, 0.01, (4, 64, 80, 2)).astype(np.float32)
    npm[1] = 1.0
    npm[2, 5:7, 5:8] = 1.0   # 6 px: removed (<= 9)
    npm[3, 10:30, 10:40] = 1.0
    inst, table = _run(pp, npm, hvm, None)
    for i in range(4):
        ref = po.proc_np_hv(npm[i], hvm[i])
        assert np.array_equal(inst[i], ref), i
    assert table["count"][0] == 0 and table["count"][2] == 0
    assert (table["type"][3][: table["count"][3]] == -1).all()


def test_large_tile_and_batch(pp):
    """1024-px post-processing tile (pretrained_model.yaml tile_shape) and a batch > SM count: same result as per-image runs."""
    npm, hvm, tpm = hovernet_head_maps(1, 512, seed=3)
    inst, _ = _run(pp, npm, hvm, tpm)
    assert np.array_equal(inst[0], po.proc_np_hv(npm[0], hvm[0]))
    npm, hvm, tpm = hovernet_head_maps(4, 164, seed=9)
    big = [np.concatenate([a] * 40) for a in (npm, hvm, tpm)]
    inst, table = _run(pp, *big)
    single, _ = _run(pp, npm, hvm, tpm)
    for r in range(40):
        assert np.array_equal(inst[4 * r: 4 * r + 4], single)
