"""8-bit RGB <-> Lab and ReinhardNormalizer without a GPU.

* the device's per-pixel conversion code (csrc/lab_math.cuh, compiled for the host behind tiab_dbg_lab_convert) against cv2 on ALL
  2^24 inputs, both directions — the same guard the Macenko tissue mask has for the L channel;
* the committed table header against its generator;
* the Reinhard oracle (oracle/reinhard_oracle.py) against the outputs of the unmodified reference class (tests/golden/reinhard_golden.npz)
  and, in the build container, against the reference class itself;
* the device's statistics + per-level maps (tiab_dbg_reinhard_maps) against numpy / cv2, and the complete device data flow emulated
  on the CPU (histograms -> maps -> map lookup between the two conversions) against the reference outputs."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import cv2
import numpy as np
import pytest

from oracle import ref_shim
from oracle import reinhard_oracle as ro
from oracle.make_golden_reinhard import inputs

GOLD = np.load(Path(__file__).parent / "golden" / "reinhard_golden.npz")


def _convert(lib, arr: np.ndarray, inverse: int) -> np.ndarray:
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    out = np.empty_like(arr)
    assert lib.tiab_dbg_lab_convert(arr.ctypes.data, out.ctypes.data, arr.size // 3, inverse) == 0
    return out


def _maps(lib, hist: np.ndarray, npix: int, t_means, t_stds):
    hist = np.ascontiguousarray(hist, dtype=np.uint32)
    tm, ts = np.ascontiguousarray(t_means, dtype=np.float64), np.ascontiguousarray(t_stds, dtype=np.float64)
    stats, lut = np.zeros(6, np.float64), np.zeros(768, np.uint8)
    dp = C.POINTER(C.c_double)
    rc = lib.tiab_dbg_reinhard_maps(hist.ctypes.data, float(npix), tm.ctypes.data_as(dp), ts.ctypes.data_as(dp), stats.ctypes.data_as(dp),
                                    lut.ctypes.data)
    return rc, stats, lut.reshape(3, 256)


def test_device_lab_conversion_equals_cv2_on_all_inputs(built_lib):
    g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    bad_fwd = bad_inv = 0
    for r in range(256):
        img = np.stack([np.full_like(g, r), g, b], axis=-1)
        bad_fwd += int((_convert(built_lib, img, 0) != cv2.cvtColor(img, cv2.COLOR_RGB2LAB)).any(-1).sum())
        bad_inv += int((_convert(built_lib, img, 1) != cv2.cvtColor(img, cv2.COLOR_LAB2RGB)).any(-1).sum())
    assert bad_fwd == 0 and bad_inv == 0, (bad_fwd, bad_inv)


def test_table_header_is_what_the_generator_writes():
    import importlib.util

    path = Path(__file__).resolve().parents[1] / "tiatoolbox_b200" / "csrc"
    spec = importlib.util.spec_from_file_location("gen_lab_tables", path / "gen_lab_tables.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = (path / "lab_tables.h").read_text()
    for name, arr in (("kSrgbGamma2040", gen.srgb_gamma_tab()), ("kLabCbrt", gen.lab_cbrt_tab()), ("kLabToYF", gen.lab_to_yf_tab()),
                      ("kSrgbInvGamma", gen.srgb_inv_gamma_tab())):
        body = text.split(f"{name}[{len(arr)}] = {{")[1].split("};")[0]
        assert np.array_equal(np.array([int(v) for v in body.replace("\n", " ").split(",")]), arr.astype(np.int64)), name


def test_oracle_equals_reference_outputs():
    target, tiles, odd, noise = inputs()
    means, stds = ro.get_mean_std(target)
    assert np.array_equal(np.array(means), GOLD["target_means"]) and np.array_equal(np.array(stds), GOLD["target_stds"])
    for i, t in enumerate(tiles):
        assert np.array_equal(ro.transform(t, means, stds), GOLD["tiles_out"][i]), i
    assert np.array_equal(ro.transform(odd, means, stds), GOLD["odd_out"])
    assert np.array_equal(ro.transform(noise, means, stds), GOLD["noise_out"])
    assert np.array_equal(ro.transform(tiles[1], tuple(float(v) for v in GOLD["png_means"]), tuple(float(v) for v in GOLD["png_stds"])), GOLD["png_tile_out"])
    m, s = ro.get_mean_std(tiles[0])
    assert np.array_equal(np.array(m), GOLD["tile0_means"]) and np.array_equal(np.array(s), GOLD["tile0_stds"])


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present (GPU box)")
def test_oracle_equals_running_reference():
    from tiatoolbox_b200.synthetic import he_target, he_tiles

    sn = ref_shim.load_stain()
    norm = sn.get_normalizer("reinhard")
    target = he_target(120, 90, seed=5)
    norm.fit(target.copy())
    means, stds = ro.get_mean_std(target)
    assert means == norm.target_means and stds == norm.target_stds
    for t in he_tiles(3, 50, 70, seed=9):
        assert np.array_equal(ro.transform(t, means, stds), norm.transform(t.copy()))


def _emulated_device_transform(lib, img: np.ndarray, t_means, t_stds):
    """histograms of the 8-bit Lab channels -> tiab_dbg_reinhard_maps -> lookup -> Lab2RGB, all through the device's own host-compiled code"""
    lab = _convert(lib, img, 0)
    hist = np.stack([np.bincount(lab[..., c].ravel(), minlength=256) for c in range(3)])
    rc, stats, lut = _maps(lib, hist, lab.shape[0] * lab.shape[1], t_means, t_stds)
    out = np.stack([lut[c][lab[..., c]] for c in range(3)], axis=-1)
    return rc, stats, lut, _convert(lib, out, 1)


def test_device_statistics_and_level_maps(built_lib):
    target, tiles, odd, noise = inputs()
    t_means, t_stds = tuple(float(v) for v in GOLD["target_means"]), tuple(float(v) for v in GOLD["target_stds"])   # Python floats
    for img in [*tiles, odd, noise, target]:
        rc, stats, lut, _ = _emulated_device_transform(built_lib, img, t_means, t_stds)
        assert rc == 0
        means, stds = ro.get_mean_std(img)
        np.testing.assert_allclose(stats[:3], means, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(stats[3:], stds, rtol=1e-12, atol=1e-13)
        # the maps are exactly the reference arithmetic on the 256 levels, given the device's own statistics
        assert np.array_equal(lut, ro.level_maps(tuple(float(v) for v in stats[:3]), tuple(float(v) for v in stats[3:]), t_means, t_stds))


def test_emulated_device_flow_equals_reference_outputs(built_lib):
    target, tiles, odd, noise = inputs()
    t_means, t_stds = tuple(float(v) for v in GOLD["target_means"]), tuple(float(v) for v in GOLD["target_stds"])   # Python floats
    cases = [(tiles[0], GOLD["tiles_out"][0]), (tiles[1], GOLD["tiles_out"][1]), (tiles[2], GOLD["tiles_out"][2]), (odd, GOLD["odd_out"]),
             (noise, GOLD["noise_out"])]
    for k, (img, want) in enumerate(cases):
        _, _, _, got = _emulated_device_transform(built_lib, img, t_means, t_stds)
        assert np.array_equal(got, want), k
    _, _, _, got = _emulated_device_transform(built_lib, tiles[1], tuple(float(v) for v in GOLD["png_means"]), tuple(float(v) for v in GOLD["png_stds"]))
    assert np.array_equal(got, GOLD["png_tile_out"])


def test_constant_channel_is_flagged(built_lib):
    img = np.full((16, 16, 3), 200, np.uint8)
    rc, stats, lut, out = _emulated_device_transform(built_lib, img, (50.0, 1.0, 2.0), (10.0, 3.0, 4.0))
    assert rc == 1 and np.all(stats[3:] == 0) and np.array_equal(lut, np.tile(np.arange(256, dtype=np.uint8), (3, 1)))
    grey = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2).reshape(16, 16, 3)   # a = b = 128 everywhere
    assert _emulated_device_transform(built_lib, grey, (50.0, 1.0, 2.0), (10.0, 3.0, 4.0))[0] == 1
