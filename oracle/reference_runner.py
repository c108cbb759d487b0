"""Run the UNMODIFIED reference implementation of the hot path (CPU or its own CUDA path) — BASELINE INFRASTRUCTURE ONLY.

Used by `bench.py --impl reference`, by `bench.py`'s `cpu_baseline` / `library_baseline` legs and by tests; never by the product
(see oracle/__init__.py).  The reference modules are loaded by oracle/ref_shim.py from /root/reference (build container) or from the
byte-for-byte vendored copy oracle/_ref (GPU box; oracle/vendor_ref.py).  What runs is the reference's own code:

  stain      tiatoolbox/tools/stainnorm.py   {Macenko,Vahadane,Ruifrok,Reinhard}Normalizer.fit / .transform       (:68-113, 222-367)
  networks   tiatoolbox/models/architecture/{hovernet,unet,vanilla}.py   <Model>.infer_batch(model, batch, device=...)  (hovernet.py:862-910)
  post-proc  HoVerNet.postproc (hovernet.py:751-859) — with the two scikit-image functions it imports (remove_small_objects, watershed:
             scikit-image is a pinned dependency that this image does not have) provided by oracle/hovernet_postproc_oracle.py, compiled C
             variant (oracle/csrc/postproc_oracle.c) so that the CPU arm is not handicapped by a Python heap.

Weights are the seeded + BN-calibrated test weights (pretrained checkpoints are remote); the reference modules load them unchanged because
the product keeps the reference's state-dict keys.
"""
from __future__ import annotations

import os

import numpy as np

_STATE = {}


def reference_available() -> bool:
    from oracle import ref_shim

    return ref_shim.available()


def normalizer(method: str, target: np.ndarray):
    """A fitted reference `StainNormalizer` (stainnorm.py:370-425 `get_normalizer`)."""
    from oracle import ref_shim

    sn = ref_shim.load_stain()
    norm = sn.get_normalizer(method)
    norm.fit(target.copy())
    return norm


def models():
    """Reference architecture modules with the scikit-image restatements installed for HoVerNet.postproc."""
    if "models" not in _STATE:
        import sys

        from oracle import hovernet_postproc_oracle as po
        from oracle import ref_shim

        m = ref_shim.load_models()
        m.hovernet.remove_small_objects = po.remove_small_objects_c
        m.hovernet.watershed = lambda image, markers=None, mask=None: po.watershed_c(image, markers, mask)
        sys.modules["skimage.morphology"].remove_small_objects = po.remove_small_objects_c
        _STATE["models"] = m
    return _STATE["models"]


def hovernet(sd, mode: str = "fast", num_types: int = 6):
    m = models()
    net = m.hovernet.HoVerNet(num_types=num_types, mode=mode).eval()
    net.load_state_dict(sd)
    return net, m.hovernet.HoVerNet


def resnet18(sd, num_classes: int = 9):
    m = models()
    net = m.vanilla.CNNModel("resnet18", num_classes).eval()
    net.load_state_dict(sd)
    return net, m.vanilla.CNNModel


def unet(sd):
    m = models()
    net = m.unet.UNetModel(3, 2, "resnet50", decoder_block=[3]).eval()
    net.load_state_dict(sd)
    return net, m.unet.UNetModel


# ---- process-pool workers (fork: the fitted normaliser / model object is inherited through _STATE) ----
def _stain_worker(chunk: np.ndarray) -> np.ndarray:
    norm = _STATE["norm"]
    return np.stack([norm.transform(t.copy()) for t in chunk])


def _postproc_worker(args):
    net = _STATE["hovernet_for_postproc"]
    out = []
    for np_map, hv_map, tp_map in args:
        res = net.postproc([np_map, hv_map, tp_map])[0]
        out.append((res["predictions"], len(res["info_dict"]["box"])))
    return out


class CpuPipeline:
    """stain transform (process pool) -> infer_batch on the CPU (torch threads) -> HoVerNet.postproc (process pool), all reference code."""

    def __init__(self, method: str | None, target: np.ndarray | None, procs: int):
        import multiprocessing as mp

        self.procs = max(1, procs)
        if method is not None:
            _STATE["norm"] = normalizer(method, target)
        self.method = method
        self.pool = None
        self._ctx = mp.get_context("fork")

    def start_pool(self):
        """Fork the workers (call after every object they need has been put into _STATE, before torch spins up its thread pool)."""
        if self.procs > 1 and self.pool is None:
            self.pool = self._ctx.Pool(self.procs)

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None

    def _map(self, fn, chunks):
        chunks = [c for c in chunks if len(c)]
        if self.pool is None or len(chunks) <= 1:
            return [fn(c) for c in chunks]
        return self.pool.map(fn, chunks)

    def stain(self, tiles: np.ndarray) -> np.ndarray:
        if self.method is None:
            return tiles
        return np.concatenate(self._map(_stain_worker, np.array_split(tiles, min(self.procs, len(tiles)))))

    def postproc(self, heads) -> list:
        items = list(zip(*heads))
        k = min(self.procs, len(items))
        chunks = [items[i::k] for i in range(k)]
        return [r for part in self._map(_postproc_worker, chunks) for r in part]


def set_postproc_model(net) -> None:
    _STATE["hovernet_for_postproc"] = net


def best_torch_threads(fn, cores: int) -> int:
    """oneDNN convolutions on many-core hosts slow down when every hyper-thread is used: give the CPU arm its best setting."""
    import time

    import torch

    best, best_t = cores, float("inf")
    for th in sorted({min(cores, t) for t in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(th)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = th, dt
    torch.set_num_threads(best)
    return best


def host_cores() -> int:
    return os.cpu_count() or 1
