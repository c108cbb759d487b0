"""Generate tests/golden/augment_golden.npz from the UNMODIFIED reference StainAugmentor (tiatoolbox/tools/stainaugment.py, loaded through
oracle/ref_shim.load_stainaugment) with its random draws replaced by fixed (alpha, beta) pairs.
Run in the build container (needs /root/reference):  python oracle/make_golden_augment.py"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from tiatoolbox_b200.synthetic import he_tiles  # noqa: E402

STAIN_MATRIX = np.array([[0.91633014, -0.20408072, -0.34451435], [0.17669817, 0.92528011, 0.33561059]])   # the example matrix of stainaugment.py:72-73
DRAWS = [((1.3, 0.1), (0.7, -0.15)), ((0.62, -0.2), (1.38, 0.2)), ((1.0, 0.0), (1.0, 0.0))]                 # ((alpha_H, beta_H), (alpha_E, beta_E))


def fixed_params(aug, pairs):
    it = iter(pairs)

    def get_params():
        aug.alpha, aug.beta = next(it)
        return {}

    aug.get_params = get_params


def main():
    sa = ref_shim.load_stainaugment()
    img = he_tiles(1, 96, 128, seed=31)[0]
    data = {"img": img}
    for k, pairs in enumerate(DRAWS):
        for bg in (False, True):
            aug = sa.StainAugmentor(method="macenko", stain_matrix=STAIN_MATRIX.copy(), augment_background=bg)
            aug.fit(img.copy(), threshold=0.85)
            fixed_params(aug, pairs)
            data[f"out{k}_{'bg' if bg else 'tissue'}"] = aug.augment()
    aug = sa.StainAugmentor(method="macenko", augment_background=False)      # matrix extracted from the image itself
    aug.fit(img.copy())
    fixed_params(aug, DRAWS[0])
    data["macenko_matrix"] = np.asarray(aug.stain_matrix)
    data["macenko_out"] = aug.augment()
    out = ROOT / "tests" / "golden" / "augment_golden.npz"
    np.savez_compressed(out, **data)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
