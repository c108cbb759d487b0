"""Generate tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_shim.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden [stain|cnn|postproc|all]

The fixtures are small and committed; the GPU box has no /root/reference, so the
``-m gpu`` tests compare the CUDA path with these files and with the oracle restatements.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def make_stain() -> None:
    from oracle import ref_shim
    from tiatoolbox_b200.synthetic import he_target, he_tiles

    sn = ref_shim.load_stain()
    misc = ref_shim.load_misc()
    out: dict[str, np.ndarray] = {}

    # the reference's own golden vector (tests/test_utils.py:882-911), re-run through the reference
    ce_in = np.array(
        [
            [[37, 244, 193], [106, 235, 128], [71, 140, 47]],
            [[103, 184, 72], [20, 188, 238], [126, 7, 0]],
            [[137, 195, 204], [32, 203, 170], [101, 77, 133]],
        ],
        dtype=np.uint8,
    )
    out["contrast_in"] = ce_in
    out["contrast_out"] = misc.contrast_enhancer(ce_in, low_p=2, high_p=98)

    real_target = ref_shim.target_image()
    syn_target = he_target(512, 512, seed=1234)
    tiles256 = he_tiles(2, 256, 256, seed=0)
    tiles64 = he_tiles(4, 64, 64, seed=7)
    tiles_odd = he_tiles(2, 100, 75, seed=11)
    sets = {"t256": tiles256, "t64": tiles64, "todd": tiles_odd}
    for method in ("macenko", "ruifrok", "vahadane", "custom"):
        kw = {}
        if method == "custom":
            kw["stain_matrix"] = np.array([[0.60, 0.72, 0.34], [0.10, 0.95, 0.30]])
            out["custom_matrix"] = kw["stain_matrix"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nr = sn.get_normalizer(method, **kw)
            nr.fit(real_target.copy())
            out[f"{method}_realfit_M"] = nr.stain_matrix_target
            out[f"{method}_realfit_maxC"] = nr.maxC_target
            ns = sn.get_normalizer(method, **kw)
            ns.fit(syn_target.copy())
        out[f"{method}_synfit_M"] = ns.stain_matrix_target
        out[f"{method}_synfit_maxC"] = ns.maxC_target
        for name, tiles in sets.items():
            if name == "t256" and method in ("ruifrok", "custom"):
                continue  # keep the fixture small; the fixed-matrix methods are covered by t64/todd
            ms = []
            res = []
            for t in tiles:
                ms.append(np.array(ns.extractor.get_stain_matrix(t.copy())))
                res.append(ns.transform(t.copy()))
            out[f"{method}_{name}_Msrc"] = np.stack(ms)
            out[f"{method}_{name}_out"] = np.stack(res)
    # tissue masks for the 256 tiles (bit-exact integer work)
    out["t256_mask"] = np.stack([misc.get_luminosity_tissue_mask(t, threshold=0.8) for t in tiles256])
    np.savez_compressed(GOLDEN / "stain_golden.npz", **out)
    print("wrote", GOLDEN / "stain_golden.npz", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    GOLDEN.mkdir(parents=True, exist_ok=True)
    if what in ("stain", "all"):
        make_stain()
    if what in ("cnn", "all"):
        from oracle.make_golden_cnn import make_cnn  # noqa: PLC0415

        make_cnn()
    if what in ("postproc", "all"):
        from oracle.make_golden_cnn import make_postproc  # noqa: PLC0415

        make_postproc()
