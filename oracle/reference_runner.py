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
def _worker_init() -> None:
    """One BLAS / OpenCV / torch thread per worker process: the pool already uses every core, and 128 processes x 128 BLAS threads
    would otherwise make the reference look ~30x slower than it is (measured on the GPU box in r2h: 0.7 s instead of 20 ms per tile)."""
    try:
        import threadpoolctl

        _STATE["tp_limit"] = threadpoolctl.threadpool_limits(1)
    except Exception:  # noqa: BLE001, S110
        pass
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:  # noqa: BLE001, S110
        pass
    try:
        import torch

        torch.set_num_threads(1)
    except Exception:  # noqa: BLE001, S110
        pass


def _stain_worker(chunk: np.ndarray) -> np.ndarray:
    norm = _STATE["norm"]
    return np.stack([norm.transform(t.copy()) for t in chunk])


def _forward_init(kind: str, threads: int) -> None:
    """Forward worker (SPAWNED, not forked: a forked child of a process that has already run torch / CUDA work can hang in OpenMP):
    builds its own copy of the reference module with the seeded + calibrated test weights."""
    import torch

    from tests import cnn_common as cc

    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    if kind == "hovernet":
        net, cls = hovernet(cc.hovernet_test_weights("fast", 6))
    elif kind == "resnet18":
        net, cls = resnet18(cc.resnet18_test_weights(9))
    else:
        net, cls = unet(cc.unet_test_weights())
    _STATE["forward_model"] = (cls, net)


def _forward_worker(batch: np.ndarray):
    """`<Model>.infer_batch(model, batch, device="cpu")` (the reference's own static method) on one batch."""
    import torch

    cls, net = _STATE["forward_model"]
    return cls.infer_batch(net, torch.from_numpy(batch), device="cpu")


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

    def start_pool(self, forward_model=None, cores: int | None = None, threads_per_forward: int = 16, kind: str = "hovernet"):
        """Fork the stain / post-processing workers (numpy / OpenCV only) and, with `forward_model` = (model class, module), SPAWN
        cores // threads_per_forward forward workers with `threads_per_forward` torch threads each, so that the CPU arm's network
        forward uses every core too (oneDNN convolutions stop scaling at ~16 threads on these hosts: r2d measured 16 as the best
        single-process setting on the 128-thread GPU box).  With fewer than 2 x threads_per_forward cores the forward runs in this
        process on all cores."""
        if self.procs > 1 and self.pool is None:
            self.pool = self._ctx.Pool(self.procs, initializer=_worker_init)
        self.fwd_pool, self.fwd_workers, self.fwd_threads = None, 1, cores or os.cpu_count() or 1
        if forward_model is not None:
            import multiprocessing as mp

            _STATE["forward_model"] = forward_model
            cores = cores or os.cpu_count() or 1
            workers = max(1, cores // threads_per_forward)
            if workers > 1:
                self.fwd_pool = mp.get_context("spawn").Pool(workers, initializer=_forward_init, initargs=(kind, threads_per_forward))
                self.fwd_workers, self.fwd_threads = workers, threads_per_forward

    def forward(self, batches: list) -> list:
        """`infer_batch` over a list of numpy batches: on the forward workers when there are several, else in this process."""
        if self.fwd_pool is not None and len(batches) > 1:
            return self.fwd_pool.map_async(_forward_worker, batches).get(timeout=900)
        import torch

        torch.set_num_threads(self.fwd_threads)
        return [_forward_worker(b) for b in batches]

    def close(self):
        for pool in (self.pool, getattr(self, "fwd_pool", None)):
            if pool is not None:
                pool.terminate()
        self.pool = self.fwd_pool = None

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
