"""Generate tests/golden/reinhard_golden.npz from the UNMODIFIED reference ReinhardNormalizer (tiatoolbox/tools/stainnorm.py:222-367,
loaded through oracle/ref_shim.load_stain): target statistics and transformed tiles for seeded synthetic H&E tiles.
Run in the build container (needs /root/reference):  python oracle/make_golden_reinhard.py"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from tiatoolbox_b200.synthetic import he_target, he_tiles  # noqa: E402


def inputs():
    target = he_target(160, 200, seed=77)
    tiles = he_tiles(3, 96, 128, seed=21)
    odd = he_tiles(1, 67, 45, seed=22)[0]          # H*W*3 not a multiple of 4: the kernels' scalar path
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)   # saturates the clip on every channel
    return target, tiles, odd, noise


def main():
    sn = ref_shim.load_stain()
    target, tiles, odd, noise = inputs()
    norm = sn.get_normalizer("reinhard")
    norm.fit(target.copy())
    data = {"target_means": np.array(norm.target_means), "target_stds": np.array(norm.target_stds)}
    data["tiles_out"] = np.stack([norm.transform(t.copy()) for t in tiles])
    data["odd_out"] = norm.transform(odd.copy())
    data["noise_out"] = norm.transform(noise.copy())
    m, s = norm.get_mean_std(tiles[0].copy())
    data["tile0_means"], data["tile0_stds"] = np.array(m), np.array(s)
    # a second target: the shipped tiatoolbox/data/target_image.png statistics (the CLI's default target)
    big = ref_shim.target_image()
    norm.fit(big.copy())
    data["png_means"], data["png_stds"] = np.array(norm.target_means), np.array(norm.target_stds)
    data["png_tile_out"] = norm.transform(tiles[1].copy())
    out = ROOT / "tests" / "golden" / "reinhard_golden.npz"
    np.savez_compressed(out, **data)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
