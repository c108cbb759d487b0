"""WSI tile-mode bookkeeping: the oracle restatement (oracle/wsi_oracle.py) and the product (tiatoolbox_b200/models/engine/wsi_merge.py,
tools/patchextraction.py) against (a) the committed outputs of the unmodified reference functions (tests/golden/wsi_merge_golden.npz,
oracle/make_golden_wsi.py) and (b) — in the build container — the reference functions themselves through oracle/ref_shim."""
from __future__ import annotations

import copy
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_shim
from oracle import wsi_oracle as wo
from oracle.make_golden_wsi import MERGE_CASES, TILE_CASES, paste_case, ref_tile_info
from tests import wsi_common as wc
from tiatoolbox_b200.models.engine import wsi_merge as wm
from tiatoolbox_b200.tools.patchextraction import PatchExtractor

GOLD = np.load(Path(__file__).parent / "golden" / "wsi_merge_golden.npz")
needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present (GPU box)")


def _same_sets(a, b) -> bool:
    return len(a) == len(b) and all(
        np.array_equal(np.asarray(x[0]).reshape(-1, 4), np.asarray(y[0]).reshape(-1, 4))
        and np.array_equal(np.asarray(x[1]).reshape(-1, 4), np.asarray(y[1]).reshape(-1, 4)) for x, y in zip(a, b))


def _product_replay(tiles, margin):
    st = wm.WsiInstanceStore()
    for bounds, flag, mode, recs in copy.deepcopy(tiles):
        st.merge_tile(wm.tile_instance_dict(wc.pack_columns(recs)), tile_bounds=bounds, tile_flag=flag, tile_mode=mode, margin=margin)
    return st


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_tile_info_matches_reference_golden(impl):
    fn = wo.get_tile_info if impl == "oracle" else wm.get_tile_info
    for i, (shape, margin) in enumerate(TILE_CASES):
        got = fn(shape, (1024, 1024), (164, 164), margin)
        want = [[GOLD[f"tile{i}.boxes{s}"], GOLD[f"tile{i}.flags{s}"]] for s in range(int(GOLD[f"tile{i}.nsets"]))]
        assert _same_sets(got, want), (impl, shape, margin)


def test_get_coordinates_matches_reference_golden():
    kw = dict(image_shape=(1000, 777), patch_input_shape=(256, 256), patch_output_shape=(164, 164), stride_shape=(164, 164))
    ins, outs = PatchExtractor.get_coordinates(**kw)
    assert np.array_equal(ins, GOLD["coords.in"]) and np.array_equal(outs, GOLD["coords.out"])
    oi, oo = wo.get_coordinates(**kw)
    assert np.array_equal(oi, GOLD["coords.in"]) and np.array_equal(oo, GOLD["coords.out"])
    # without an output shape only the input boxes come back, and they equal the output grid
    only = PatchExtractor.get_coordinates(image_shape=(500, 300), patch_input_shape=(200, 200), stride_shape=(200, 200))
    assert only.shape == (6, 4) and np.array_equal(only[-1], [400, 200, 600, 400])
    # within-bound filtering
    i2, o2 = PatchExtractor.get_coordinates(**kw, input_within_bound=True)
    assert len(i2) and np.all(i2[:, :2] >= 0) and np.all(i2[:, 2] <= 1000) and np.all(i2[:, 3] <= 777)
    i3, o3 = PatchExtractor.get_coordinates(**kw, output_within_bound=True)
    assert np.all(o3[:, 2] <= 1000) and np.all(o3[:, 3] <= 777) and len(o3) == (1000 // 164) * (777 // 164)


def test_get_coordinates_errors_like_the_reference():
    ok = dict(image_shape=(100, 100), patch_input_shape=(32, 32), patch_output_shape=(16, 16), stride_shape=(16, 16))
    for key, bad, rx in (("image_shape", (10.5, 3), "Invalid `image_shape`"), ("patch_input_shape", (-1, 3), "Invalid `patch_input_shape`"),
                         ("patch_output_shape", (1, 2, 3), "Invalid `patch_output_shape`"), ("stride_shape", (1.0, 2), "Invalid `stride_shape`")):
        with pytest.raises(ValueError, match=rx):
            PatchExtractor.get_coordinates(**{**ok, key: bad})
    with pytest.raises(ValueError, match="must larger than"):
        PatchExtractor.get_coordinates(**{**ok, "patch_output_shape": (64, 64)})
    with pytest.raises(ValueError, match="must > 1"):
        PatchExtractor.get_coordinates(**{**ok, "stride_shape": (0, 16)})


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_merge_replay_matches_reference_golden(impl):
    for seed, shape, margin in MERGE_CASES:
        tiles = wc.scenario(seed, image_shape=shape, margin=margin, density=30)
        kept = wo.replay_merge(copy.deepcopy(tiles), margin) if impl == "oracle" else _product_replay(tiles, margin).alive_records()
        boxes = np.array([r["box"] for r in kept], dtype=np.int64).reshape(-1, 4)
        assert np.array_equal(boxes, GOLD[f"merge{seed}.boxes"]), (impl, seed)
        assert np.array_equal(np.array([r["prob"] for r in kept]), GOLD[f"merge{seed}.prob"]), (impl, seed)


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_paste_instance_tiles_matches_reference_golden(impl):
    for seed in (0, 1):
        wsi = np.zeros((260, 300), np.int32)
        mx, trace = None, []
        for bounds, pred in paste_case(seed):
            if impl == "oracle":
                mx = wo.paste_instance_tile(wsi, pred, bounds, mx)
            else:
                mx = wm.update_tile_based_predictions_array(pred, wsi, bounds, (0, 0), mx)
            trace.append(int(mx))
        assert np.array_equal(wsi, GOLD[f"paste{seed}.wsi"]), (impl, seed)
        assert trace == GOLD[f"paste{seed}.max"].tolist(), (impl, seed)


def test_store_bookkeeping_and_info_dict_format():
    tiles = wc.scenario(11, image_shape=(2500, 2150), margin=128, density=20)
    st = _product_replay(tiles, 128)
    recs = st.alive_records()
    assert len(st) == len(recs) > 0
    info = st.to_info_dict()
    assert set(info) == set(wm.INFO_KEYS) and all(v.dtype == object and len(v) == len(recs) for v in info.values())
    shifted = st.to_info_dict(offset=(7, -3))
    assert np.array_equal(shifted["box"][0], info["box"][0] + np.array([7, -3, 7, -3]))
    assert np.array_equal(shifted["contours"][0], info["contours"][0] + np.array([7, -3])) and shifted["contours"][0].dtype == np.int32
    assert np.array_equal(shifted["centroid"][0], info["centroid"][0] + np.array([7, -3]))
    assert shifted["prob"][0] == info["prob"][0]
    assert wm.WsiInstanceStore().to_info_dict() == {}          # the reference's DataFrame of nothing has no columns
    # growth past the initial capacity keeps order and evictions
    big = wm.WsiInstanceStore()
    rec = {"box": np.array([1, 1, 5, 5]), "centroid": np.array([3.0, 3.0]), "contours": np.zeros((4, 2), np.int32), "prob": 0.5, "type": 1}
    for k in range(3000):
        big.merge_tile({1: dict(rec, prob=float(k))}, tile_bounds=[10 * k, 0, 10 * k + 2000, 2000], tile_flag=[0, 0, 0, 0], tile_mode=0, margin=0)
    assert len(big) == 3000 and [r["prob"] for r in big.alive_records()][:3] == [0.0, 1.0, 2.0]
    with pytest.raises(ValueError, match="Unknown tile mode 7"):
        big.merge_tile({1: rec}, tile_bounds=[0, 0, 100, 100], tile_flag=[0, 0, 0, 0], tile_mode=7, margin=0)


@needs_ref
def test_against_the_running_reference_functions():
    """Fresh seeds (not in the golden file) through the unmodified reference code, the oracle and the product."""
    m = ref_shim.load_wsi_merge()
    for shape in [(1500, 1200), (2100, 990), (984, 2000), (2952, 2952), (100, 100)]:
        for margin in (128, 32, None):
            want = ref_tile_info(m, shape, (1024, 1024), (164, 164), margin)
            assert _same_sets(wo.get_tile_info(shape, (1024, 1024), (164, 164), margin), want)
            assert _same_sets(wm.get_tile_info(shape, (1024, 1024), (164, 164), margin), want)
    for seed in range(100, 108):
        shape = [(2500, 2150), (3100, 1200), (1100, 3000), (2952, 2952)][seed % 4]
        margin = [128, 64, 128, 200][seed % 4]
        tiles = wc.scenario(seed, image_shape=shape, margin=margin, density=40)
        want = wc.replay_reference(m, tiles, margin)
        assert wc.records_equal(wo.replay_merge(copy.deepcopy(tiles), margin), want), seed
        assert wc.records_equal(_product_replay(tiles, margin).alive_records(), want), seed
    ins, outs = m.PatchExtractor.get_coordinates(image_shape=(3333, 1111), patch_input_shape=(270, 270), patch_output_shape=(80, 80),
                                                 stride_shape=(80, 80))
    pi, po = PatchExtractor.get_coordinates(image_shape=(3333, 1111), patch_input_shape=(270, 270), patch_output_shape=(80, 80),
                                            stride_shape=(80, 80))
    assert np.array_equal(ins, pi) and np.array_equal(outs, po)


def test_column_store_equals_record_store_on_random_tile_sequences():
    """`WsiColumnStore` (vectorised per tile, what WSI mode uses) == `WsiInstanceStore` (per record, pinned to the unmodified reference
    functions above): same survivors, same order, same values and dtypes, with and without a final offset."""
    for seed in range(6):
        for margin in (128, 0, 37):
            tiles = wc.scenario(seed, margin=margin, density=25)
            want_store = _product_replay(tiles, margin)
            cs = wm.WsiColumnStore()
            consume = bool((seed + margin) & 1)       # the engine hands its tile arrays over (vertices shifted in place): same result
            for bounds, flag, mode, recs in copy.deepcopy(tiles):
                cs.merge_tile(wm.columns_from_records(recs), tile_bounds=bounds, tile_flag=flag, tile_mode=mode, margin=margin,
                              consume=consume)
            assert len(cs) == len(want_store)
            for off in ((0, 0), (5, -9)):
                got, want = cs.to_info_dict(offset=off), want_store.to_info_dict(offset=off)
                assert set(got) == set(want) or (not got and not want)
                for key in want:
                    assert len(got[key]) == len(want[key])
                    for a, b in zip(got[key], want[key]):
                        assert np.array_equal(np.asarray(a), np.asarray(b)), (seed, margin, key)
                        if key in ("box", "centroid", "contours"):
                            assert np.asarray(a).dtype == np.asarray(b).dtype, (key, np.asarray(a).dtype, np.asarray(b).dtype)
    assert wm.WsiColumnStore().to_info_dict() == {}


def test_tensor_paste_equals_the_numpy_recurrence_including_int32_wraparound():
    """`paste_instance_tile_tensor` (torch, branch-free, what WSI mode runs on the device) == `update_tile_based_predictions_array` (numpy,
    pinned to the reference function above) over a random tile sequence — also when the running `max_inst_value` passes 2^31 (the
    reference's recurrence roughly doubles it per tile, so real slides wrap after ~31 tiles)."""
    import warnings

    import torch

    rng = np.random.default_rng(5)
    for start in (None, 2**31 - 400):
        H, W = 300, 340
        want = np.zeros((H, W), np.int32)
        got = torch.zeros((H, W), dtype=torch.int32)
        m_np = None if start is None else np.int32(start)
        m_t = torch.tensor(0 if start is None else start, dtype=torch.int32)
        lut = torch.zeros(120 * 120 + 2 + wm.PASTE_SCRATCH, dtype=torch.int32)
        for _ in range(40):
            th, tw = int(rng.integers(20, 120)), int(rng.integers(20, 120))
            x0, y0 = int(rng.integers(0, W - 10)), int(rng.integers(0, H - 10))
            tile = np.zeros((th, tw), np.int32)
            for lab in rng.permutation(np.arange(1, 9))[: int(rng.integers(0, 8))]:
                yy, xx = int(rng.integers(0, th - 5)), int(rng.integers(0, tw - 5))
                tile[yy: yy + int(rng.integers(3, 30)), xx: xx + int(rng.integers(3, 30))] = lab
            bounds = (x0, y0, x0 + tw, y0 + th)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                m_np = wm.update_tile_based_predictions_array(tile, want, bounds, (0, 0), m_np)
            m_t = wm.paste_instance_tile_tensor(torch.from_numpy(tile), got, bounds, m_t, lut)
            assert int(lut.abs().sum()) == 0
            assert np.array_equal(got.numpy(), want)
            assert int(m_t) == int(np.int32(m_np))
