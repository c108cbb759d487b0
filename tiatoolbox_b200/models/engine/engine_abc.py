"""EngineABC — patch-mode orchestration on the B200.

Keeps the constructor and `run()` surface of tiatoolbox/models/engine/engine_abc.py:280-335, 1684-1829
(`predict` is kept as an alias of `run`: the 1.x name BASELINE.json remembers).  What changes is the data
flow of `_run_patch_mode` (:1374-1488): the reference does one H2D + one blocking D2H per batch and keeps
everything downstream on the CPU; here a batch goes pinned host -> device once, stain normalisation (when
the model's `preproc_func` is a B200 `StainNormalizer.transform`), the network and the post-processing all
run on the device, and only the final results come back.

Multi-GPU (one process per GPU, torchrun): patches are sharded in contiguous index blocks (row-major grid
order of tools/patchextraction.py:576-591 is preserved) with zero traffic during compute and ONE all-gather
of the results at the end (NCCL).

Scope: `patch_mode=True`, `output_type="dict"`.  WSI mode / zarr / AnnotationStore outputs
are the callers and data formats around the hot path and raise NotImplementedError here.
"""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from tiatoolbox_b200.models.architecture import get_pretrained_model
from tiatoolbox_b200.models.models_abc import ModelABC

logger = logging.getLogger(__name__)


def cast_to_min_dtype(array: np.ndarray) -> np.ndarray:
    """utils/misc.py:1925-1961: bool if max == 1 else the smallest unsigned integer type."""
    max_value = np.max(array) if array.size else 0
    if max_value == 1:
        return array.astype(bool)
    for max_allowed, dtype in ((255, np.uint8), (65535, np.uint16), (2**32 - 1, np.uint32), (2**64 - 1, np.uint64)):
        if max_value <= max_allowed:
            return array.astype(dtype)
    return array


class _PooledBlock:
    """Owner of one pooled byte block: the `.base` of the result array AND of every view of it (numpy collapses view chains to the first
    non-array owner), so the block goes back to the pool only when the last view is gone.  Python-level buffer protocol: PEP 688."""

    def __init__(self, pool: "HostResultPool", nbytes: int, raw: np.ndarray) -> None:
        self.pool, self.nbytes, self.raw = pool, nbytes, raw

    def __buffer__(self, flags: int) -> memoryview:
        return memoryview(self.raw)

    def __del__(self) -> None:
        try:
            self.pool._give_back(self.nbytes, self.raw)  # noqa: SLF001
        except Exception:  # noqa: BLE001, S110 — interpreter shutdown: the pool's module may already be torn down
            pass


class HostResultPool:
    """Recycles the host memory of large result arrays.  A result such as `predictions` (107 KB per image, for ALL images of a multi-rank
    job on every rank) is a fresh numpy array per `run()`; faulting in and zeroing fresh pages costs ~0.3 ms per MB — 66 ms per call at
    2048 images, more than 15 % of an 8-GPU run (r2s).  Arrays handed out here sit on a pooled byte block owned by a `_PooledBlock`;
    when the array and every view of it have been garbage-collected the block returns to the pool, so a steady stream of `run()` calls
    writes into pages that are already mapped.  A request takes the smallest free block of at least its size and at most 25 % more; the
    free list is bounded (`max_free_bytes`, oldest blocks dropped first), so calls of ever-changing sizes cannot pile up memory."""

    def __init__(self, max_free_bytes: int = 1 << 30) -> None:
        self.free: list = []                      # [(nbytes, raw)], oldest first
        self.max_free_bytes = int(max_free_bytes)

    def free_bytes(self) -> int:
        return sum(n for n, _ in self.free)

    def take(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(shape, dtype=np.int64)) * dtype.itemsize, 1)
        fits = [(n, i) for i, (n, _) in enumerate(self.free) if nbytes <= n <= nbytes + nbytes // 4]
        if fits:
            size, raw = self.free.pop(min(fits)[1])
        else:
            size, raw = nbytes, np.empty(nbytes, dtype=np.uint8)
        return np.ndarray(tuple(shape), dtype=dtype, buffer=_PooledBlock(self, size, raw))

    def _give_back(self, nbytes: int, raw) -> None:
        if nbytes > self.max_free_bytes:
            return
        self.free.append((nbytes, raw))
        while self.free_bytes() > self.max_free_bytes:
            self.free.pop(0)


_RESULT_POOL = HostResultPool()


class EngineABC:
    """Template: run -> shard -> batches -> model on device -> post-process -> gather -> dict."""

    def __init__(self, model: str | ModelABC, batch_size: int = 8, num_workers: int = 0, weights: str | Path | None = None, *,
                 device: str = "cuda", verbose: bool = False) -> None:
        self.batch_size = batch_size
        self.num_workers = num_workers  # kept for API compatibility: patches are fed from host arrays, no DataLoader
        self.verbose = verbose
        self.ioconfig = None
        if isinstance(model, str):
            self.model, self.ioconfig = get_pretrained_model(model, weights)
        elif isinstance(model, ModelABC):
            self.model = model
            if weights is not None:
                self.model.load_weights_from_file(weights)
        else:
            msg = "model must be a registry name or a tiatoolbox_b200 ModelABC (no CPU / generic nn.Module fallback on this path)"
            raise TypeError(msg)
        dev = torch.device(device)
        if dev.type != "cuda":
            msg = "tiatoolbox_b200 engines need a CUDA device: the hot path has no CPU fallback"
            raise ValueError(msg)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.device = dev
        self.model.to(dev)
        self.return_probabilities = False
        self.return_labels = False
        self.labels = None
        self.drop_keys: list[str] = []

    # ---- distributed helpers ----
    @staticmethod
    def _dist():
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist, dist.get_rank(), dist.get_world_size()
        return None, 0, 1

    @staticmethod
    def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
        """Contiguous block [r*ceil(n/R), (r+1)*ceil(n/R))."""
        per = -(-n // world)
        return min(n, rank * per), min(n, (rank + 1) * per)

    @staticmethod
    def _all_gather_rows(t: torch.Tensor, n_total: int) -> torch.Tensor:
        """All-gather per-patch result rows (dim 0) of every rank's contiguous shard (one collective;
        NCCL on GPUs, gloo in the CPU tests).  A rank whose shard is empty learns the row shape first."""
        dist, rank, world = EngineABC._dist()
        if dist is None:
            return t
        shape = torch.zeros(8, dtype=torch.long, device=t.device)
        shape[0] = t.dim()
        for i, d in enumerate(t.shape[1:]):
            shape[1 + i] = d if t.shape[0] else 0
        dist.all_reduce(shape, op=dist.ReduceOp.MAX)
        if t.shape[0] == 0:
            t = torch.zeros((0, *[int(v) for v in shape[1:int(shape[0])]]), dtype=t.dtype, device=t.device)
        per = -(-n_total // world)
        pad = per - t.shape[0]
        if pad:
            t = torch.cat([t, torch.zeros((pad, *t.shape[1:]), dtype=t.dtype, device=t.device)])
        out = torch.empty((world * per, *t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out[:n_total]

    # ---- input handling ----
    @staticmethod
    def _as_patch_array(images) -> np.ndarray:
        if isinstance(images, np.ndarray):
            arr = images
        elif isinstance(images, (list, tuple)) and len(images) and isinstance(images[0], np.ndarray):
            arr = np.stack(images)
        else:
            msg = "patch_mode expects a (N, H, W, 3) array or a list of HxWx3 arrays (file / WSI inputs are outside the B200 path)"
            raise NotImplementedError(msg)
        if arr.ndim != 4 or arr.shape[-1] != 3:
            msg = f"expected (N, H, W, 3) patches, got {arr.shape}"
            raise ValueError(msg)
        if arr.dtype != np.uint8:
            arr = np.clip(np.rint(arr), 0, 255).astype(np.uint8)
        return np.ascontiguousarray(arr)

    def _device_preproc(self, x: torch.Tensor) -> torch.Tensor:
        """Apply model.preproc_func on the device when it is a B200 stain normaliser's transform
        (the wiring of examples/inference-pipelines/slide-graph.ipynb); other callables run per patch on the host
        like the reference's dataset does (dataset_abc.py:439, 522)."""
        from tiatoolbox_b200.tools.stainnorm import StainNormalizer

        func = self.model.preproc_func
        owner = getattr(func, "__self__", None)
        if isinstance(owner, StainNormalizer) and getattr(func, "__name__", "") == "transform":
            # A patch whose stain fit fails (empty tissue mask on background / 255-padded patches, degenerate estimate) passes through
            # UNNORMALISED — the reference raises ValueError for such a patch (utils/misc.py:286-288).  Deliberate deviation: a slide
            # must not die on its background patches.  The failures are counted on the device (no sync per batch), reported once per
            # run() (`stain_failures`, a warning), and `raise_on_stain_failure=True` restores the reference's behaviour.
            out, info = owner.transform_batch(x, return_info=True, raise_on_error=False)
            failed = (info["status"] != 0).sum()
            acc = self.__dict__.get("_stain_failed")
            self.__dict__["_stain_failed"] = failed if acc is None else acc + failed
            return out
        if func is self.model.preproc or func is ModelABC.preproc:
            return x
        host = x.cpu().numpy()
        return torch.from_numpy(np.stack([np.asarray(func(p)) for p in host])).to(x.device)

    # ---- the public entry point ----
    def run(self, images, *, masks=None, input_resolutions=None, patch_input_shape=None, ioconfig=None,  # noqa: ARG002
            patch_mode: bool = True, save_dir=None, overwrite: bool = False, output_type: str = "dict", **kwargs):  # noqa: ARG002
        if not patch_mode:
            msg = "WSI mode (stitching / tile de-duplication) is the next row of the hot-path scope (SURVEY.md §8f), not built yet"
            raise NotImplementedError(msg)
        if output_type.lower() != "dict":
            msg = "only output_type='dict' is produced on the B200 path (zarr / AnnotationStore are result persistence)"
            raise NotImplementedError(msg)
        for k, v in kwargs.items():  # the reference sets every kwarg as an attribute (engine_abc.py:1306-1307)
            setattr(self, k, v)
        patches = self._as_patch_array(images)
        n_total = len(patches)
        dist, rank, world = self._dist()
        lo, hi = self.shard_bounds(n_total, rank, world)
        self.model._plan_batch = self.batch_size
        self._n_total = n_total
        self._stain_failed = None
        raw = self.infer_patches(patches[lo:hi])
        self._report_stain_failures()
        out = self.post_process_patches(raw, n_total)
        if self.labels is not None:
            out["labels"] = np.asarray(self.labels)
        return out

    predict = run  # tiatoolbox 1.x name

    #: raise ValueError (like the reference) instead of passing a patch through unnormalised when its stain fit fails
    raise_on_stain_failure = False

    def _report_stain_failures(self) -> int:
        acc = self.__dict__.get("_stain_failed")
        self.stain_failures = 0 if acc is None else int(acc.item())
        self.__dict__["_stain_failed"] = None
        if self.stain_failures:
            msg = (f"{self.stain_failures} patch(es) could not be stain-normalised (empty tissue mask / degenerate estimate) and were "
                   "passed to the network unnormalised")
            if self.raise_on_stain_failure:
                raise ValueError("Empty tissue mask computed. " + msg)
            logger.warning(msg)
        return self.stain_failures

    # ---- to be specialised ----
    def infer_patches(self, patches: np.ndarray):
        raise NotImplementedError

    def post_process_patches(self, raw, n_total: int) -> dict:
        raise NotImplementedError

    def _batches(self, patches: np.ndarray):
        """Yield device uint8 batches (pinned staging, async H2D) after the device-side preproc."""
        n = len(patches)
        if n == 0:
            return
        # two pinned staging buffers: the host fills one while the H2D copy of the other may still be in flight
        shape = (min(self.batch_size, n), *patches.shape[1:])
        pins = [self._pinned(("h2d", k), shape, torch.uint8) for k in range(2)]
        done = [None, None]
        for b, i in enumerate(range(0, n, self.batch_size)):
            chunk = patches[i:i + self.batch_size]
            pin = pins[b & 1]
            if done[b & 1] is not None:
                done[b & 1].synchronize()
            pin[: len(chunk)].copy_(torch.from_numpy(chunk))
            x = pin[: len(chunk)].to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            done[b & 1] = ev
            yield self._device_preproc(x)

    def _pinned(self, key, shape, dtype) -> torch.Tensor:
        """Cached pinned host buffer (cudaHostAlloc costs milliseconds; engines are called once per slide region)."""
        cache = self.__dict__.setdefault("_pin_cache", {})
        k = (key, tuple(shape), dtype)
        if k not in cache:
            cache[k] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
        return cache[k]

    def _to_host_async(self, key, t: torch.Tensor, side_stream: bool = False):
        """Start a device -> pinned-host copy; the caller synchronises on an event before reading.  Default: on the current stream.
        `side_stream=True`: on the engine's copy stream, ordered after the work already queued on the current stream, so that the copy
        overlaps the NEXT batch's kernels instead of sitting in front of them; returns (buffer, event recorded after the copy)."""
        buf = self._pinned(key, (max(self.batch_size, t.shape[0]), *t.shape[1:]), t.dtype)[: t.shape[0]]
        self.__dict__["d2h_bytes"] = self.__dict__.get("d2h_bytes", 0) + buf.numel() * buf.element_size()   # what actually crosses the bus
        if not side_stream or t.device.type != "cuda":
            buf.copy_(t, non_blocking=True)
            if side_stream:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                return buf, ev
            return buf
        cs = self.__dict__.get("_copy_stream")
        if cs is None:
            cs = self.__dict__["_copy_stream"] = torch.cuda.Stream(self.device)
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        cs.wait_event(ready)
        with torch.cuda.stream(cs):
            buf.copy_(t, non_blocking=True)
            t.record_stream(cs)          # the caching allocator must not hand `t`'s memory out before the copy has read it
            ev = torch.cuda.Event()
            ev.record(cs)
        return buf, ev
