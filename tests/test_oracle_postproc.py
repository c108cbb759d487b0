"""CPU tests pinning the post-processing oracle: the two restated scikit-image functions on hand-built
known-answer cases + properties, and (when /root/reference is present) the restated pipeline against the
reference's own unmodified `_proc_np_hv` / `get_instance_info` with those two functions injected."""
from __future__ import annotations

import sys

import numpy as np
import pytest
from scipy import ndimage

from oracle import hovernet_postproc_oracle as po
from tiatoolbox_b200.synthetic import hovernet_head_maps


def test_remove_small_objects_known_answer():
    lab = np.array([[1, 1, 0, 2], [1, 0, 0, 2], [0, 3, 3, 3], [0, 3, 3, 0]], np.int32)
    out = po.remove_small_objects(lab, max_size=2)
    assert np.array_equal(out, np.array([[1, 1, 0, 0], [1, 0, 0, 0], [0, 3, 3, 3], [0, 3, 3, 0]], np.int32))
    assert np.array_equal(po.remove_small_objects(lab, max_size=3), np.where(lab == 3, 3, 0))
    assert np.array_equal(lab, np.array([[1, 1, 0, 2], [1, 0, 0, 2], [0, 3, 3, 3], [0, 3, 3, 0]], np.int32))  # input untouched


def test_watershed_known_answers():
    # 1-D valley pair: the ridge pixel goes to the marker whose front reaches it first (lower neighbour popped first)
    img = np.array([[0.0, 1.0, 2.0, 5.0, 3.0, 1.5, 0.5]])
    markers = np.array([[1, 0, 0, 0, 0, 0, 2]])
    mask = np.ones_like(markers, bool)
    out = po.watershed(img, markers, mask)
    assert out.tolist() == [[1, 1, 1, 1, 2, 2, 2]]
    # mask blocks flooding; markers outside the mask are dropped; unreachable pixels stay 0
    img = np.zeros((3, 5))
    markers = np.zeros((3, 5), int)
    markers[1, 0] = 7
    markers[0, 4] = 9
    mask = np.ones((3, 5), bool)
    mask[:, 2] = False
    mask[0, 4] = False
    out = po.watershed(img, markers, mask)
    assert (out[:, :2] == 7).all() and (out[:, 2:] == 0).all()
    # plateau: insertion age gives breadth-first growth, neighbour order (-row, -1, +1, +row)
    img = np.zeros((1, 6))
    markers = np.array([[1, 0, 0, 0, 0, 2]])
    out = po.watershed(img, markers, np.ones((1, 6), bool))
    assert out.tolist() == [[1, 1, 1, 2, 2, 2]]


def test_watershed_properties():
    npm, hvm, _ = hovernet_head_maps(2, 164, seed=1)
    for i in range(2):
        dbg = {}
        out = po.proc_np_hv(npm[i], hvm[i], debug=dbg)
        blb, marker = dbg["blb"], dbg["marker"] * dbg["blb"]
        assert set(np.unique(out)) <= set(np.unique(marker))           # labels come from markers
        assert (out[blb == 0] == 0).all()                               # mask respected
        assert (out[marker > 0] == marker[marker > 0]).all()            # markers keep their label
        assert np.array_equal(po.watershed(dbg["dist"], out, blb), out)  # idempotent on a labelled input


def test_pipeline_equals_reference_code_path(have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import ref_shim

    m = ref_shim.load_models()
    m.hovernet.remove_small_objects = po.remove_small_objects
    m.hovernet.watershed = po.watershed
    npm, hvm, tpm = hovernet_head_maps(2, 164, seed=2)
    for i in range(2):
        ref = m.hovernet.HoVerNet._proc_np_hv(npm[i], hvm[i])
        assert np.array_equal(po.proc_np_hv(npm[i], hvm[i]), ref)
        tp = np.around(tpm[i]).astype("uint8")
        a = m.hovernet.HoVerNet.get_instance_info(ref, tp, offset=(5, 9), verbose=False)
        b = po.get_instance_info(ref, tp, offset=(5, 9))
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert np.array_equal(a[k]["box"], b[k]["box"]) and np.array_equal(a[k]["contours"], b[k]["contours"])
            assert np.allclose(a[k]["centroid"], b[k]["centroid"], atol=0) and a[k]["type"] == b[k]["type"] and a[k]["prob"] == b[k]["prob"]


def test_c_restatement_equals_python_restatement_on_adversarial_floods():
    """oracle/csrc/postproc_oracle.c (independently written binary-heap flood) == the Python heapq restatement on plateaus, equal-valued
    markers, single-pixel bridges and random quantised landscapes (many (value) ties), and on the pipeline's own intermediates."""
    rng = np.random.default_rng(7)
    cases = []
    h, w = 24, 31
    flat = np.zeros((h, w))                                   # one big plateau, two markers: pure (age, index) ordering
    mk = np.zeros((h, w), np.int32)
    mk[3, 4], mk[20, 25] = 1, 2
    cases.append((flat, mk, np.ones((h, w), bool)))
    mk2 = np.zeros((h, w), np.int32)                          # equal-valued marker blobs next to each other
    mk2[5:8, 5:8], mk2[5:8, 9:12], mk2[15:18, 20:23] = 1, 2, 3
    cases.append((np.round(rng.random((h, w)) * 3) / 3, mk2, np.ones((h, w), bool)))
    bridge = np.zeros((h, w), bool)                           # two rooms joined by a one-pixel corridor
    bridge[2:12, 2:12] = bridge[2:12, 16:28] = True
    bridge[6, 12:16] = True
    mk3 = np.zeros((h, w), np.int32)
    mk3[4, 4], mk3[9, 25] = 5, 9
    cases.append((rng.random((h, w)), mk3, bridge))
    for _ in range(40):                                       # random quantised landscapes: ties everywhere
        hh, ww = int(rng.integers(5, 40)), int(rng.integers(5, 40))
        img = np.round(rng.normal(size=(hh, ww)) * rng.integers(1, 4)) / 2
        msk = rng.random((hh, ww)) < 0.8
        m = np.zeros((hh, ww), np.int32)
        for lab in range(1, int(rng.integers(2, 7))):
            y, x = int(rng.integers(0, hh)), int(rng.integers(0, ww))
            m[max(0, y - 1): y + 2, max(0, x - 1): x + 2] = lab
        cases.append((img, m, msk))
    for img, m, msk in cases:
        assert np.array_equal(po.watershed_c(img, m, msk), po.watershed(img, m, msk))
    npm, hvm, _ = hovernet_head_maps(3, 164, seed=11)
    for i in range(3):
        dbg = {}
        out = po.proc_np_hv(npm[i], hvm[i], debug=dbg)
        assert np.array_equal(po.watershed_c(dbg["dist"], dbg["marker"], dbg["blb"]), out)
        lab = ndimage.label(npm[i][..., 0] >= 0.5)[0]
        assert np.array_equal(po.remove_small_objects_c(lab, 9), po.remove_small_objects(lab, 9))
