"""Device contour tracing (tiab_hovernet_contours) against the oracle / cv2 on synthetic head maps, and the engine's opt-in
`device_contours` mode against its default host-cv2 mode."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import hovernet_postproc_oracle as po
from tiatoolbox_b200.synthetic import hovernet_head_maps

pytestmark = pytest.mark.gpu


def test_device_contours_equal_cv2(built_lib):
    import torch

    from tiatoolbox_b200.models import postproc as pp

    npm, hvm, tpm = hovernet_head_maps(4, 164, seed=12)
    inst, table = pp.proc_np_hv_batch(torch.from_numpy(npm).cuda(), torch.from_numpy(hvm).cuda(), torch.from_numpy(tpm).cuda())
    pp.trace_contours(inst, table)
    inst_h = inst.cpu().numpy()
    rows = {k: v.cpu().numpy() for k, v in table.items()}
    assert (rows["cont_total"] > 100).all() and (rows["cont_total"] <= rows["cont_xy"].shape[1]).all()
    for i in range(4):
        row = {k: v[i] for k, v in rows.items()}
        got = pp.instance_info(inst_h[i], row, offset=(5, 2))
        host_only = {k: v for k, v in row.items() if not k.startswith("cont_")}
        want = pp.instance_info(inst_h[i], host_only, offset=(5, 2))               # cv2.findContours per instance
        ref = po.get_instance_info(inst_h[i], np.around(tpm[i]).astype("uint8"), offset=(5, 2))
        assert list(got) == list(want) == list(ref) and len(got) > 10
        for k in got:
            assert got[k]["contours"].dtype == np.int32
            assert np.array_equal(got[k]["contours"], want[k]["contours"]) and np.array_equal(got[k]["contours"], ref[k]["contours"])
            assert np.array_equal(got[k]["box"], want[k]["box"])
    # a tiny vertex buffer: the affected image falls back to the host tracer and still gives the same answer
    inst2, table2 = pp.proc_np_hv_batch(torch.from_numpy(npm).cuda(), torch.from_numpy(hvm).cuda(), torch.from_numpy(tpm).cuda())
    pp.trace_contours(inst2, table2, capacity=64)
    row = {k: v[0].cpu().numpy() for k, v in table2.items()}
    assert int(row["cont_total"]) > 64
    small = pp.instance_info(inst_h[0], row)
    full = pp.instance_info(inst_h[0], {k: v[0] for k, v in rows.items()})
    assert list(small) == list(full) and all(np.array_equal(small[k]["contours"], full[k]["contours"]) for k in full)


def test_engine_device_contours_equal_default(built_lib):
    from tests.cnn_common import hovernet_test_weights
    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_tiles

    eng = NucleusInstanceSegmentor("hovernet_fast-pannuke", batch_size=2, device="cuda")
    eng.model.load_state_dict(hovernet_test_weights("fast", 6))
    tiles = he_tiles(5, 256, 256, seed=0)
    a = eng.run(tiles)
    eng.device_contours = True
    b = eng.run(tiles)
    assert np.array_equal(a["predictions"], b["predictions"])
    for i in range(5):
        assert len(a["contours"][i]) == len(b["contours"][i])
        for ca, cb in zip(a["contours"][i], b["contours"][i]):
            assert np.array_equal(ca, cb)
        for ba, bb in zip(a["box"][i], b["box"][i]):
            assert np.array_equal(ba, bb)
