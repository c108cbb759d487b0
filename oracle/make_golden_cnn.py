"""CNN fixtures: (i) calibrated BatchNorm statistics for the seeded test weights, (ii) outputs of the
UNMODIFIED reference modules (loaded through oracle/ref_shim.py) for those weights on seeded tiles.

TEST INFRASTRUCTURE ONLY; run here (needs /root/reference):  python -m oracle.make_golden cnn
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _f16(a: torch.Tensor) -> np.ndarray:
    return a.detach().cpu().numpy().astype(np.float32)


def make_hovernet() -> None:
    from oracle import ref_shim
    from tiatoolbox_b200.models.architecture._params import calibrate, seeded_state_dict
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_forward_reference, hovernet_spec
    from tiatoolbox_b200.synthetic import he_tiles

    m = ref_shim.load_models()
    out = {}
    for tag, mode, nt, size in (("fast", "fast", 6, 256),):
        sd = seeded_state_dict(hovernet_spec(3, nt, mode), seed=0)
        xcal = torch.from_numpy(he_tiles(2, size, size, seed=100)).permute(0, 3, 1, 2).float()
        stats = calibrate(sd, lambda s, x: hovernet_forward_reference(s, x, mode, nt), xcal)
        np.savez_compressed(GOLDEN / f"hovernet_{tag}_bn_calib.npz", **{k: v.numpy() for k, v in stats.items()})
        net = m.hovernet.HoVerNet(num_types=nt, mode=mode).eval()
        net.load_state_dict(sd)
        tiles = he_tiles(2, size, size, seed=0)
        with torch.inference_mode():
            raw = net(torch.from_numpy(tiles).permute(0, 3, 1, 2).float())
        heads = m.hovernet.HoVerNet.infer_batch(net, torch.from_numpy(tiles), device="cpu")
        for k, v in raw.items():  # raw branch outputs, every 2nd pixel (keeps the fixture ~1 MB)
            out[f"{tag}_logits_{k}"] = _f16(v.permute(0, 2, 3, 1))[:, ::2, ::2]
        for k, v in zip(("np", "hv", "tp"), heads):  # infer_batch heads of tile 0, full resolution
            out[f"{tag}_head_{k}"] = np.asarray(v, dtype=np.float32)[:1]
    np.savez_compressed(GOLDEN / "hovernet_golden.npz", **out)
    print("hovernet golden:", {k: v.shape for k, v in out.items()})


def make_cnn() -> None:
    GOLDEN.mkdir(parents=True, exist_ok=True)
    make_hovernet()


def make_postproc() -> None:
    pass
