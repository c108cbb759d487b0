#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the stain + patch-CNN hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
                    [--workload macenko_hovernet|macenko|resnet18|fcn_tissue_mask|wsi]

Headline metric (BASELINE.json, default workload `macenko_hovernet`): patches/sec on 256x256x3 tiles for Macenko stain
normalisation + HoVer-Net (hovernet_fast-pannuke architecture) incl. the on-device watershed post-processing, instance
table and contours.  One "step" = one pass of that path over one batch of synthetic tiles per GPU.
The other workloads are BASELINE.json's configs 1, 2, 3 and 5 (`--workload ...`), one JSON line each.

Head maps.  Pretrained checkpoints are remote, and the heads of a random-weight HoVer-Net contain almost no nuclei, which
would make the post-processing a no-op.  Following SURVEY.md §8d config 4 the network's forward runs in full on every tile
and the post-processing (watershed, instance table, contours, D2H, host assembly) is fed SURVEY's synthetic head maps
(K ~ U{20..60} nuclei per 164-px map) instead of the network's own: `BenchHoVerNet.infer_batch_device` below.  This holds
for `value`, for `e2e` and for the reference arm alike, and is stated in `config.heads`.

At N>1 the driver launches this file under torchrun, one rank per GPU; tiles are sharded across ranks (weak scaling:
per-GPU batch fixed), no data-path collective except the all-gather of the final instance maps.  Rank 0 prints ONE JSON line.

`value`             device-timed whole-job patches/s, inputs resident in HBM (CUDA events, max over ranks)
`e2e`               the same metric through the public engine API (NucleusInstanceSegmentor.run with the stain normaliser installed
                    as preproc_func) with HOST numpy buffers: H2D + kernels + D2H + host assembly of the result dict
`roofline`          the dominant kernel (conv_gemm_kernel, tensor-bound): ALGORITHMIC FLOPs (297.87 GFLOP per
                    patch) / its summed launch durations (CUDA events per launch, measured live) vs MEASURED_PEAKS.json;
                    `by_class` splits the launches into tensor-bound and HBM-bound ones, each against its own roof
`roofline_stain`    the fused stain kernel (HBM-bound) measured the same way on a 16 384-tile launch
`library_baseline`  the reference's OWN CUDA path (its unmodified nn.Module, eager cuDNN, TF32 on / off) on the same GPU and tiles
`cpu_baseline`      the unmodified reference's CPU path on the box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PATCH = 256
STAIN_BYTES_PER_PATCH = 2 * PATCH * PATCH * 3      # read u8 once + write u8 once
GFLOP = {"hovernet_fast": 297.87, "resnet18": 3.627, "fcn_tissue_mask": 291.15}   # SURVEY.md Appendix C (2*MAC of all convs / linear)
METRICS = {
    "macenko_hovernet": "patches/sec (256x256x3) Macenko + HoVer-Net (hovernet_fast) incl. on-device post-processing",
    "macenko": "patches/sec (256x256x3) Macenko stain-normalisation transform",
    "resnet18": "patches/sec (224x224x3) PatchPredictor resnet18-kather100k",
    "fcn_tissue_mask": "tiles/sec (1024x1024x3) SemanticSegmentor fcn-tissue_mask",
    "wsi": "patches/sec (256x256x3) Vahadane stain-norm + HoVer-Net end to end on a synthetic 20k-patch WSI",
}
UNITS = {"fcn_tissue_mask": "tiles/s"}


def common_config(workload: str) -> dict:
    """The part of `config` both arms share (the driver compares it)."""
    cfg = {
        "macenko_hovernet": {"workload": "macenko_hovernet", "tile": [PATCH, PATCH, 3], "stain": "macenko, fit on a synthetic 256x256 target",
                             "network": "hovernet_fast (num_types=6), seeded + BN-calibrated random weights (residual-branch gain 0.2-0.25)",
                             "l2_policy": "GPU arm: 11.9 GB of activations per 16-tile step >> 126 MB L2, no flush needed",
                             "heads": "network forward runs in full; post-processing input = SURVEY 8d config-4 synthetic head maps "
                                      "(K~U{20..60} nuclei per 164x164 map), because random-weight heads contain no nuclei"},
        "macenko": {"workload": "macenko", "tile": [PATCH, PATCH, 3], "stain": "macenko, fit on a synthetic 256x256 target",
                    "l2_policy": "GPU arm: 8192 tiles = 1.6 GB in + 1.6 GB out per step >> 126 MB L2, no flush needed"},
        "resnet18": {"workload": "resnet18", "tile": [224, 224, 3], "network": "resnet18-kather100k architecture (9 classes), seeded + calibrated weights",
                     "l2_policy": "GPU arm: 1024 tiles per step, activations >> 126 MB L2, no flush needed"},
        "fcn_tissue_mask": {"workload": "fcn_tissue_mask", "tile": [1024, 1024, 3],
                            "network": "fcn-tissue_mask architecture (UNet, ResNet-50 encoder), seeded + calibrated weights",
                            "l2_policy": "GPU arm: 4 tiles of 1024x1024 per step, activations >> 126 MB L2, no flush needed"},
        "wsi": {"workload": "wsi", "tile": [PATCH, PATCH, 3], "stain": "vahadane, fit on a synthetic 256x256 target",
                "network": "hovernet_fast (num_types=6), seeded + BN-calibrated random weights",
                "heads": "synthetic config-4 head maps per patch (as in macenko_hovernet)",
                "l2_policy": "one pass over a 1.6 GB slide, nothing is re-read",
                "slide": "virtual slide, 141 x 142 output patches of 164 px (SURVEY 8d config 5)"},
    }
    return dict(cfg[workload])


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]), "tf_sustained": float(d["bf16_tflops_sustained"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# synthetic inputs and seeded weights (shared by both arms)
# ---------------------------------------------------------------------------------------------------
def _tiles(n_unique: int, seed: int = 0, size: int = PATCH) -> np.ndarray:
    from tiatoolbox_b200.synthetic import he_tiles

    return he_tiles(n_unique, size, size, seed=seed)


def _calibrated(spec, fname):
    from tiatoolbox_b200.models.architecture._params import seeded_state_dict, with_calibration

    return with_calibration(seeded_state_dict(spec, seed=0), ROOT / "tests" / "golden" / fname)


def _hovernet_weights():
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec

    return _calibrated(hovernet_spec(3, 6, "fast"), "hovernet_fast_bn_calib.npz")


def _resnet18_weights():
    from tiatoolbox_b200.models.architecture.vanilla import cnn_spec

    return _calibrated(cnn_spec("resnet18", 9), "resnet18_bn_calib.npz")


def _unet_weights():
    from tiatoolbox_b200.models.architecture.unet import unet_spec

    return _calibrated(unet_spec(3, 2, (3,)), "unet_bn_calib.npz")


def _head_pool(n: int, seed: int = 0):
    from tiatoolbox_b200.synthetic import hovernet_head_maps

    return hovernet_head_maps(n, 164, seed=seed)


def synthetic_slide(n_patches: int, seed: int = 0) -> np.ndarray:
    """H&E-like slide whose 164-px output grid has about `n_patches` cells (141 x 142 for 20 000), built from seeded 256-px tiles."""
    import math

    nx = max(1, int(math.floor(math.sqrt(n_patches))))
    ny = max(1, int(math.ceil(n_patches / nx)))
    h, w = ny * 164, nx * 164
    base = _tiles(16, seed=seed)
    rng = np.random.default_rng(seed)
    slide = np.empty((h, w, 3), np.uint8)
    for y in range(0, h, 256):
        for x in range(0, w, 256):
            t = base[int(rng.integers(0, len(base)))]
            slide[y:y + 256, x:x + 256] = t[: min(256, h - y), : min(256, w - x)]
    return slide


def make_bench_hovernet(pool_maps: int, device):
    """HoVerNet subclass for the bench: the REAL plan runs on every tile; the head maps handed to the post-processing are the synthetic
    config-4 maps (rotating through a device-resident pool), because the random-weight heads hold no nuclei (module docstring)."""
    import torch

    from tiatoolbox_b200.models.architecture.hovernet import HoVerNet

    class BenchHoVerNet(HoVerNet):
        _pool = None
        _cursor = 0

        @staticmethod
        def infer_batch_device(model, batch_data, *, device):
            real = HoVerNet.infer_batch_device(model, batch_data, device=device)     # full forward, heads included
            n = real[0].shape[0]
            pool, m = BenchHoVerNet._pool, BenchHoVerNet._pool[0].shape[0]
            idx = (torch.arange(n, device=pool[0].device) + BenchHoVerNet._cursor) % m
            BenchHoVerNet._cursor = (BenchHoVerNet._cursor + n) % m
            return tuple(p.index_select(0, idx) for p in pool)

    model = BenchHoVerNet(num_types=6, mode="fast")
    model.load_state_dict(_hovernet_weights())
    BenchHoVerNet._pool = [torch.from_numpy(a).to(device) for a in _head_pool(pool_maps, seed=0)]
    return model


# ---------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference on the host cores (oracle/reference_runner.py -> oracle/ref_shim.py -> /root/reference or oracle/_ref)
# ---------------------------------------------------------------------------------------------------
def reference_pipeline(workload: str, sample: int, cores: int):
    """Returns (step_fn() -> seconds split dict, cpu_baseline description, pool owner).  One call of step_fn processes `sample` units."""
    import torch

    from oracle import reference_runner as rr
    from tiatoolbox_b200.synthetic import he_target

    if not rr.reference_available():
        raise RuntimeError("reference sources not found (neither /root/reference nor oracle/_ref)")
    torch.set_grad_enabled(False)
    procs = min(cores, sample)
    if workload in ("macenko_hovernet", "wsi"):
        method = "macenko" if workload == "macenko_hovernet" else "vahadane"
        pipe = rr.CpuPipeline(method, he_target(PATCH, PATCH), procs)
        net, cls = rr.hovernet(_hovernet_weights())
        rr.set_postproc_model(net)
        pipe.start_pool((cls, net), cores, kind="hovernet")
        tiles = _tiles(sample, seed=0)
        heads = _head_pool(sample, seed=0)

        def step():
            t0 = time.perf_counter()
            normed = pipe.stain(tiles)
            t1 = time.perf_counter()
            fb = 2 if pipe.fwd_workers > 1 else 4
            pipe.forward([normed[i:i + fb] for i in range(0, sample, fb)])   # HoVerNet.infer_batch(model, batch, device="cpu") (hovernet.py:862-910)
            t2 = time.perf_counter()
            res = pipe.postproc(heads)                      # HoVerNet.postproc per patch (hovernet.py:751-859) on the synthetic head maps
            t3 = time.perf_counter()
            assert sum(r[1] for r in res) > 10 * sample
            return {"stain_s": t1 - t0, "forward_s": t2 - t1, "postproc_s": t3 - t2}

        desc = (f"{sample} seeded tiles per step: reference {method.capitalize()}Normalizer.transform over {procs} single-threaded processes; "
                f"reference HoVerNet.infer_batch(device='cpu') in batches of {2 if pipe.fwd_workers > 1 else 4} over {pipe.fwd_workers} worker process(es) x {pipe.fwd_threads} torch "
                f"threads; reference HoVerNet.postproc over {procs} processes (scikit-image's remove_small_objects / watershed are absent "
                "from the image: compiled-C restatements)")
        return step, desc, pipe
    if workload == "macenko":
        pipe = rr.CpuPipeline("macenko", he_target(PATCH, PATCH), procs)
        pipe.start_pool()
        tiles = _tiles(sample, seed=0)

        def step():
            t0 = time.perf_counter()
            pipe.stain(tiles)
            return {"stain_s": time.perf_counter() - t0}

        return step, f"{sample} seeded tiles per step: reference MacenkoNormalizer.transform over {procs} single-threaded processes", pipe
    if workload == "resnet18":
        net, cls = rr.resnet18(_resnet18_weights())
        pipe = rr.CpuPipeline(None, None, 1)
        pipe.start_pool((cls, net), cores, kind="resnet18")
        x = _tiles(sample, seed=0, size=224).astype(np.float32) / 255.0
        bs = max(16, sample // max(1, pipe.fwd_workers))

        def step():
            t0 = time.perf_counter()
            pipe.forward([x[i:i + bs] for i in range(0, sample, bs)])
            return {"forward_s": time.perf_counter() - t0}

        return step, (f"{sample} seeded 224x224 tiles per step: reference CNNModel.infer_batch(device='cpu') in batches of {bs} over "
                      f"{pipe.fwd_workers} worker process(es) x {pipe.fwd_threads} torch threads"), pipe
    if workload == "fcn_tissue_mask":
        net, cls = rr.unet(_unet_weights())
        pipe = rr.CpuPipeline(None, None, 1)
        pipe.start_pool((cls, net), cores, kind="unet")
        x = _tiles(sample, seed=0, size=1024)

        def step():
            t0 = time.perf_counter()
            pipe.forward([x[i:i + 1] for i in range(sample)])
            return {"forward_s": time.perf_counter() - t0}

        return step, (f"{sample} seeded 1024x1024 tiles per step: reference UNetModel.infer_batch(device='cpu'), one tile per call over "
                      f"{pipe.fwd_workers} worker process(es) x {pipe.fwd_threads} torch threads"), pipe
    raise ValueError(workload)


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = args.workload
    nfw = max(1, cores // 16)                 # forward workers of the CPU arm (oracle/reference_runner.CpuPipeline.start_pool)
    # a step = a bounded sample: two tiles per forward worker keep every core busy and (warmup + steps) x ~12 s within a few minutes
    sample = {"macenko_hovernet": max(4, 2 * nfw), "wsi": max(4, 2 * nfw), "macenko": max(64, 8 * cores), "resnet18": 64 * nfw, "fcn_tissue_mask": max(2, nfw)}[wl]
    unit = UNITS.get(wl, "patches/s")
    try:
        step, desc, pipe = reference_pipeline(wl, sample, cores)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": repr(e)[:200]}), flush=True)
        return
    total_s, n, parts = 0.0, 0, {}
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        parts = step()
        if i >= args.warmup:
            total_s += time.perf_counter() - t0
            n += sample
    if pipe is not None:
        pipe.close()
    value = n / total_s
    line = {
        "impl": "reference", "metric": METRICS[wl], "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic", "config": common_config(wl),
        "detail": {"sample_units_per_step": sample, "last_step_split_s": parts},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------------
def _event_time(fn, warm: int, iters: int) -> float:
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def conv_roofline(plan, algorithmic_flops: float, pk: dict) -> tuple[dict, float]:
    """Per-launch CUDA events over the plan's ops -> the roofline object of the conv engine (+ aux kernel ms)."""
    import torch

    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows = []
    aux_ms = 0.0
    for rep in range(3):                                      # repetition 0 warms the eager launch path
        for i, op in enumerate(plan.ops):
            a.record()
            op()
            b.record()
            torch.cuda.synchronize()
            dt = a.elapsed_time(b)
            if rep == 0:
                rows.append([op, []])
            rows[i][1].append(dt)
    conv_ms = flops_grid = mma = 0.0
    classes = {"tensor": [0, 0.0, 0.0, 0.0], "hbm": [0, 0.0, 0.0, 0.0]}    # launches, ms, flops, bytes
    for op, ts in rows:
        dt = statistics.median(ts[1:])
        if not hasattr(op, "desc"):
            aux_ms += dt
            continue
        conv_ms += dt
        flops_grid += op.flops
        mma += op.mma_flops
        t_tensor = op.flops / (pk["tf_sustained"] * 1e12)
        t_hbm = getattr(op, "bytes", 0.0) / (pk["hbm"] * 1e9)
        c = classes["tensor" if t_tensor >= t_hbm else "hbm"]
        c[0] += 1
        c[1] += dt
        c[2] += op.flops
        c[3] += getattr(op, "bytes", 0.0)
    ach = algorithmic_flops / (conv_ms / 1e3) / 1e12
    traffic = None      # dram read + write of all conv launches of one step, from the committed ncu --set full capture (HoVer-Net, 16 tiles)
    tp = ROOT / "profiles" / "conv_gemm_traffic.json"
    if tp.exists() and abs(algorithmic_flops / (GFLOP["hovernet_fast"] * 1e9) - round(algorithmic_flops / (GFLOP["hovernet_fast"] * 1e9))) < 1e-6:
        tj = json.loads(tp.read_text())
        traffic = tj["dram_bytes_per_step"] * (algorithmic_flops / (GFLOP["hovernet_fast"] * 1e9)) / tj["tiles_per_step"]
    by_class = {}
    ct = classes["tensor"]
    if ct[0]:
        tf = ct[2] / (ct[1] / 1e3) / 1e12
        by_class["tensor_bound_layers"] = {"launches": ct[0], "ms": ct[1], "achieved_tflops_on_grid_flops": tf, "frac_of_tensor_peak": tf / pk["tf_sustained"]}
    ch = classes["hbm"]
    if ch[0]:
        gb = ch[3] / (ch[1] / 1e3) / 1e9
        by_class["hbm_bound_layers"] = {"launches": ch[0], "ms": ch[1], "achieved_gbs_on_min_traffic": gb, "frac_of_hbm_peak": gb / pk["hbm"]}
    roof = {
        "bound": "tensor", "kernel": "conv_gemm_kernel (all launches of one step)", "achieved": ach, "peak": pk["tf_sustained"],
        "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained",
        "algorithmic_flops_per_step": algorithmic_flops, "grid_flops_per_step": flops_grid, "kernel_ms_per_step": conv_ms,
        "mma_issue_rate_tflops": mma / (conv_ms / 1e3) / 1e12,
        "launches_per_step": sum(1 for op, _ in rows if hasattr(op, "desc")), "by_class": by_class,
        "note": "frac is on ALGORITHMIC flops (SURVEY App. C); fp16x3 issues 3 tensor-core passes per algorithmic FLOP (cap 1/3); "
                "by_class: a launch is tensor-bound when flops/tensor-peak >= min-traffic/HBM-peak, else HBM-bound",
    }
    return roof, aux_ms


def stain_roofline(norm, dev, pk: dict, tiles: int = 16384) -> dict:
    """The fused Macenko kernel on a launch big enough to fill the GPU (16 384 tiles = 3.2 GB in + 3.2 GB out)."""
    import torch

    uniq = torch.from_numpy(_tiles(64, seed=7)).to(dev)
    x = uniq.repeat(tiles // len(uniq), 1, 1, 1).contiguous()
    ms = _event_time(lambda: norm.transform_batch(x, raise_on_error=False), 2, 5)
    ach = tiles * STAIN_BYTES_PER_PATCH / (ms / 1e3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "stain_fused_traffic.json"
    if tp.exists():
        t = json.loads(tp.read_text()).get("dram_bytes_per_tile")
        traffic = None if t is None else t * tiles
    del x
    return {"bound": "hbm", "kernel": "stain_fused_kernel", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
            "traffic": traffic, "peak_source": pk["src"], "algorithmic_bytes_per_launch": tiles * STAIN_BYTES_PER_PATCH, "kernel_ms": ms,
            "tiles_per_launch": tiles, "patches_per_s": tiles / ms * 1e3}


def library_baseline(dev, x_u8, ours_forward_ms: float) -> dict:
    """The reference's own CUDA path: its unmodified HoVerNet nn.Module (eager PyTorch / cuDNN) on the same GPU and the same tiles
    (hovernet.py:896 `model(x)` inside infer_batch), TF32 on (PyTorch's cuDNN default) and off.  Informational: the kernel to beat."""
    import torch

    from oracle import reference_runner as rr

    out = {"what": "reference HoVerNet nn.Module forward on cuda (eager PyTorch/cuDNN), same tiles, CUDA events",
           "ours_forward_ms": ours_forward_ms, "batch": int(x_u8.shape[0])}
    try:
        net, cls = rr.hovernet(_hovernet_weights())
        net = net.to(dev)
        xin = x_u8.permute(0, 3, 1, 2).float().contiguous()
        for name, tf32 in (("tf32", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.inference_mode():
                ms = _event_time(lambda: net(xin), 2, 4)
            out[f"forward_ms_{name}"] = ms
            out[f"patches_per_s_{name}"] = x_u8.shape[0] / ms * 1e3
        torch.backends.cudnn.allow_tf32 = True
        with torch.inference_mode():                      # the full public call: infer_batch incl. head activations and D2H
            t0 = time.perf_counter()
            for _ in range(2):
                cls.infer_batch(net, x_u8, device=str(dev))
            torch.cuda.synchronize()
            out["infer_batch_ms_tf32_with_d2h"] = (time.perf_counter() - t0) / 2 * 1e3
        del net
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 - an informational figure must never cost the bench line
        out["error"] = repr(e)[:200]
    return out


def cpu_baseline_leg(workload: str) -> dict:
    """Bounded sample of the reference's CPU path on this box's host cores (rank 0, N = 1 only)."""
    cores = os.cpu_count() or 1
    nfw = max(1, cores // 16)
    sample = {"macenko_hovernet": max(4, 2 * nfw), "wsi": max(4, 2 * nfw), "macenko": max(64, 4 * cores), "resnet18": 64 * nfw, "fcn_tissue_mask": max(2, nfw)}[workload]
    try:
        step, desc, pipe = reference_pipeline(workload, sample, cores)
        step()
        t0 = time.perf_counter()
        parts = step()
        dt = time.perf_counter() - t0
        if pipe is not None:
            pipe.close()
        return {"value": sample / dt, "unit": UNITS.get(workload, "patches/s"), "cores": cores, "kind": "reference", "sample": desc, "split_s": parts}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNITS.get(workload, "patches/s"), "cores": cores, "kind": "reference", "sample": f"failed: {e!r}"[:300]}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="macenko_hovernet", choices=sorted(METRICS))
    ap.add_argument("--batch", type=int, default=None, help="units per GPU per step (device-resident leg)")
    ap.add_argument("--e2e-batch", type=int, default=None, help="units per GPU per step (host-buffer leg)")
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--patches", type=int, default=20000, help="wsi workload: output patches of the virtual slide")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--stain-roofline-tiles", type=int, default=16384)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.build()
    from tiatoolbox_b200 import _lib
    from tiatoolbox_b200.models import postproc as pp
    from tiatoolbox_b200.synthetic import he_target
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload
    lib = _lib.lib()
    h = _lib.handle(local_rank)
    pk = _peaks()
    unit = UNITS.get(wl, "patches/s")
    if wl == "wsi":
        run_wsi(args, rank, world, local_rank, dev)
        return

    B = args.batch or {"macenko_hovernet": 16, "macenko": 8192, "resnet18": 1024, "fcn_tissue_mask": 4}[wl]
    # units per e2e call: a slide region feeds thousands of tiles per run(); a few engine batches are enough for the engine's host / device
    # pipelining to reach steady state while the default bench still finishes in seconds
    Be = args.e2e_batch or {"macenko_hovernet": 256, "macenko": 4096, "resnet18": 4096, "fcn_tissue_mask": 16}[wl]
    size = {"resnet18": 224, "fcn_tissue_mask": 1024}.get(wl, PATCH)
    uniq = torch.from_numpy(_tiles(min(128, max(B, 16)) if size != 1024 else 8, seed=rank, size=size)).to(dev)
    x = uniq.repeat((B + len(uniq) - 1) // len(uniq), 1, 1, 1)[:B].contiguous()

    norm = None
    if wl in ("macenko_hovernet", "macenko"):
        norm = get_normalizer("macenko")
        norm.device = local_rank
        norm.fit(he_target(PATCH, PATCH))

    plan = model = eng = None
    launches = {}
    if wl == "macenko_hovernet":
        from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor

        model = make_bench_hovernet(max(4 * B, 64), dev)
        model.precision = args.precision
        model.to(dev)
        model._plan_batch = B
        plan = model.plan_for(B, dev, (PATCH, PATCH))
        cls = type(model)

        def step():
            out, info = norm.transform_batch(x, return_info=True, raise_on_error=False)
            launches["stain"] = lib.tiab_last_launch_count(h)
            heads = cls.infer_batch_device(model, out, device=dev)
            inst, table = pp.proc_np_hv_batch(*heads)
            launches["postproc"] = lib.tiab_last_launch_count(h)
            pp.trace_contours(inst, table)
            launches["contours"] = lib.tiab_last_launch_count(h)
            if world > 1:  # the only collective: all-gather of the final result map
                gathered = torch.empty((world * inst.shape[0], *inst.shape[1:]), dtype=inst.dtype, device=dev)
                dist.all_gather_into_tensor(gathered, inst.contiguous())
            return info, table
    elif wl == "macenko":
        def step():
            _, info = norm.transform_batch(x, return_info=True, raise_on_error=False)
            launches["stain"] = lib.tiab_last_launch_count(h)
            return info, None
    elif wl == "resnet18":
        from tiatoolbox_b200.models.architecture.vanilla import CNNModel

        model = CNNModel("resnet18", 9)
        model.load_state_dict(_resnet18_weights())
        model.precision = args.precision
        model.to(dev)
        model._plan_batch = B
        plan = model.plan_for(B, dev, (224, 224))

        def step():
            probs = CNNModel.infer_batch_device(model, x, device=dev)
            if world > 1:
                g = torch.empty((world * probs.shape[0], probs.shape[1]), dtype=probs.dtype, device=dev)
                dist.all_gather_into_tensor(g, probs.contiguous())
            return None, None
    else:  # fcn_tissue_mask
        from tiatoolbox_b200.models.architecture.unet import UNetModel

        model = UNetModel(3, 2, "resnet50", decoder_block=[3])
        model.load_state_dict(_unet_weights())
        model.precision = args.precision
        model.to(dev)
        model._plan_batch = B
        plan = model.plan_for(B, dev, (1024, 1024))

        def step():
            probs = UNetModel.infer_batch_device(model, x, device=dev)
            pred = probs.argmax(-1).to(torch.uint8)
            if world > 1:
                g = torch.empty((world * pred.shape[0], *pred.shape[1:]), dtype=pred.dtype, device=dev)
                dist.all_gather_into_tensor(g, pred.contiguous())
            return None, None

    info = table = None
    for _ in range(args.warmup):
        info, table = step()
    torch.cuda.synchronize()
    if info is not None:
        assert int(info["status"].sum().item()) == 0, "synthetic tiles must all normalise"
    nuclei_per_step = int(table["count"].sum().item()) if table is not None else None
    launches_per_step = sum(launches.values()) + (plan.num_launches if plan else 0)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms / 1e3)

    # ---- per-kernel timings for the rooflines (CUDA events around every launch; rank 0 only) ----
    roofline = roofline_stain = None
    breakdown = {}
    if rank == 0:
        if norm is not None:
            breakdown["stain"] = _event_time(lambda: norm.transform_batch(x, raise_on_error=False), 1, 5)
        if plan is not None:
            key = {"macenko_hovernet": "hovernet_fast", "resnet18": "resnet18", "fcn_tissue_mask": "fcn_tissue_mask"}[wl]
            roofline, aux_ms = conv_roofline(plan, GFLOP[key] * 1e9 * B, pk)
            breakdown["conv_gemm"] = roofline["kernel_ms_per_step"]
            breakdown["aux_glue"] = aux_ms
            xin = norm.transform_batch(x, raise_on_error=False) if norm is not None else x
            breakdown["forward_graph_replay"] = _event_time(lambda: plan.forward(xin), 2, 5)
        if wl == "macenko_hovernet":
            heads = [p[:B].contiguous() for p in type(model)._pool]
            breakdown["postproc"] = _event_time(lambda: pp.proc_np_hv_batch(*heads), 2, 5)
            inst, tab = pp.proc_np_hv_batch(*heads)
            breakdown["contours"] = _event_time(lambda: pp.trace_contours(inst, tab), 2, 5)
            breakdown["note"] = ("per-launch CUDA events, eager launches; postproc / contours on the synthetic config-4 head maps the timed step uses "
                                 f"({nuclei_per_step} nuclei per step of {B} maps)")
        if norm is not None:
            roofline_stain = stain_roofline(norm, dev, pk, args.stain_roofline_tiles if wl == "macenko_hovernet" else max(B, 8192))

    # ---- e2e leg: numpy host buffers through the public API ----
    # under torchrun every rank hands the SAME list to eng.run, which shards it — so the list holds Be units per rank (weak scaling)
    sharded = wl != "macenko"
    Be_total = Be * world if sharded else Be
    host_tiles = np.ascontiguousarray(x[:min(Be_total, B)].cpu().numpy())
    if len(host_tiles) < Be_total:
        host_tiles = np.concatenate([host_tiles] * (-(-Be_total // len(host_tiles))))[:Be_total]
    if wl == "macenko_hovernet":
        eng = NucleusInstanceSegmentor(model, batch_size=B, device=str(dev))
        model.preproc_func = norm.transform
        e2e_call = lambda: eng.run(host_tiles)  # noqa: E731
        api = "NucleusInstanceSegmentor.run(numpy tiles), preproc_func = MacenkoNormalizer.transform"
    elif wl == "macenko":
        e2e_call = lambda: norm.transform_batch(host_tiles, raise_on_error=False)  # noqa: E731
        api = "StainNormalizer.transform_batch(numpy)"
    elif wl == "resnet18":
        from tiatoolbox_b200.models.engine import PatchPredictor

        eng = PatchPredictor(model, batch_size=B, device=str(dev))
        e2e_call = lambda: eng.run(host_tiles, return_probabilities=True)  # noqa: E731
        api = "PatchPredictor.run(numpy tiles, return_probabilities=True)"
    else:
        from tiatoolbox_b200.models.engine import SemanticSegmentor

        eng = SemanticSegmentor(model, batch_size=B, device=str(dev))
        e2e_call = lambda: eng.run(host_tiles)  # noqa: E731
        api = "SemanticSegmentor.run(numpy tiles)"
    r = None
    for _ in range(2):
        r = e2e_call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_units = Be_total if sharded else world * Be
    e2e_value = e2e_units * args.steps / float(te.item())
    if wl == "macenko_hovernet":
        d2h = int(getattr(eng, "d2h_bytes", Be * 164 * 164 * 4))
        e2e_nuclei = int(sum(len(b) for b in r["box"]))
    elif wl == "macenko":
        d2h, e2e_nuclei = world * (Be * PATCH * PATCH * 3 + Be * 68), None
    elif wl == "resnet18":
        d2h, e2e_nuclei = e2e_units * 9 * 4, None
    else:
        d2h, e2e_nuclei = e2e_units * 512 * 512 * 2 * 4, None

    if rank == 0:
        cfg = common_config(wl)
        line = {
            "metric": METRICS[wl], "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("u8 in / f32 math" if wl == "macenko" else
                      "f16x3 (split-fp16 operands, fp32 accumulate)" if args.precision == "fp16x3" else "f16"),
            "data": "synthetic", "config": cfg,
            "detail": {"units_per_gpu_per_step": B, "precision": args.precision if plan is not None else "f32",
                       "nuclei_per_step": nuclei_per_step,
                       "l2_policy": ("activations %.1f GB per step >> 126 MB L2" % (plan.bytes_allocated() / 1e9)) if plan is not None else
                                    f"inputs+outputs {2 * B * PATCH * PATCH * 3 / 1e6:.0f} MB per step > 126 MB L2",
                       "parallelism": f"unit-shard x{world}, all-gather of the result"},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": e2e_units * size * size * 3, "d2h_bytes_per_step": d2h,
                    "units_per_step": e2e_units, "api": api, "nuclei_per_step": e2e_nuclei,
                    "rank0_host_split_of_last_call_s": getattr(eng, "last_timings", None) if eng is not None else None},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if roofline is not None:
            line["roofline"] = roofline
            if roofline_stain is not None:
                line["roofline_stain"] = roofline_stain
        elif roofline_stain is not None:
            line["roofline"] = roofline_stain
        line["breakdown_ms"] = breakdown
        if wl == "macenko_hovernet" and world == 1 and not args.no_library_baseline:
            line["library_baseline"] = library_baseline(dev, x, breakdown.get("forward_graph_replay"))
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_wsi(args, rank, world, local_rank, dev):
    """BASELINE config 5: Vahadane -> HoVer-Net on a virtual 20k-patch slide through the WSI-mode public call; the slide is cut into bands of
    post-processing tile rows (one per rank, strong scaling), NCCL gather of the tile instance maps.  Wall clock between barriers, max over ranks."""
    import torch
    import torch.distributed as dist

    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_target
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    B = args.batch or 16
    slide = synthetic_slide(args.patches)
    norm = get_normalizer("vahadane")
    norm.device = local_rank
    norm.fit(he_target(PATCH, PATCH))
    model = make_bench_hovernet(256, dev)
    model.precision = args.precision
    eng = NucleusInstanceSegmentor(model, batch_size=B, device=str(dev))
    from tiatoolbox_b200.models.architecture import PRETRAINED_INFO

    eng.ioconfig = PRETRAINED_INFO["hovernet_fast-pannuke"]["ioconfig"]     # pretrained_model.yaml:642-672 (patch 256 -> 164, tiles 1024, margin 128)
    model.preproc_func = norm.transform

    def once(img):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.run(img, patch_mode=False, return_predictions=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter() - t0, out

    once(synthetic_slide(400))                      # warm-up: plan build, CUDA graph capture, pinned buffers
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    times, out = [], None
    for _ in range(max(1, min(args.steps, 2))):
        dt, out = once(slide)
        times.append(dt)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = len(out["coordinates"])
        v = n / float(t.item())
        line = {
            "metric": METRICS["wsi"], "value": v, "unit": "patches/s", "n_gpus": world, "steps": len(times), "warmup": 1,
            "ms_per_step": float(t.item()) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16x3 (split-fp16 operands, fp32 accumulate)", "data": "synthetic", "config": common_config("wsi"),
            "detail": {"slide_shape": list(slide.shape), "patches": n, "batch": B, "instances": int(len(out.get("box", []))),
                       "rank0_split_s": getattr(eng, "wsi_timings", None),
                       "timed": "host slide -> merged instance table + int32 WSI instance map (wall clock incl. H2D, stitching, tile post-processing, gather, merge)"},
            "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": int(slide.nbytes), "d2h_bytes_per_step": int(getattr(eng, "d2h_bytes", 0)),
                    "api": "NucleusInstanceSegmentor.run(slide, patch_mode=False, return_predictions=True), preproc_func = VahadaneNormalizer.transform"},
            "gpu_launches": None, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg("wsi")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
