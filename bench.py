#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the stain + patch-CNN hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic 256x256x3 uint8 tiles
(SURVEY.md §8d generator, seeded).  At N>1 the driver launches this file under torchrun, one rank
per GPU; tiles are sharded across ranks (weak scaling: per-GPU batch fixed) with no data-path
collective except the final all-gather of the per-tile results.  Rank 0 prints ONE JSON line.

`value`     device-timed whole-job patches/s with inputs resident in HBM (CUDA events, max over ranks)
`e2e`       the same metric through the public API with HOST buffers (H2D + kernels + D2H timed)
`roofline`  the dominant kernel's algorithmic bytes / its measured duration vs MEASURED_PEAKS.json
`cpu_baseline` the numpy oracle (a port of the reference's CPU path) timed on this box's host cores
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
BYTES_PER_PATCH_ALGO = 2 * PATCH * PATCH * 3  # read u8 once + write u8 once
METRIC = "patches/sec (256x256x3) Macenko stain-normalisation transform"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def _tiles(n_unique: int, seed: int = 0) -> np.ndarray:
    from tiatoolbox_b200.synthetic import he_tiles

    return he_tiles(n_unique, PATCH, PATCH, seed=seed)


def _oracle_worker(args):
    import numpy as np  # noqa: F811

    from oracle import stain_oracle as so

    tiles, mt, mc = args
    return [so.transform("macenko", t.copy(), mt, mc).sum() for t in tiles]


def cpu_baseline(sample: int, procs: int):
    """Time the numpy oracle (a restatement of tiatoolbox's CPU Macenko transform) on host cores."""
    from oracle import stain_oracle as so
    from tiatoolbox_b200.synthetic import he_target

    tiles = _tiles(sample, seed=0)
    mt, mc = so.fit("macenko", he_target(PATCH, PATCH))
    if procs <= 1:
        t0 = time.perf_counter()
        _oracle_worker((tiles, mt, mc))
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        chunks = [c for c in np.array_split(tiles, procs) if len(c)]
        with mp.get_context("fork").Pool(len(chunks)) as pool:
            pool.map(_oracle_worker, [(c[:1], mt, mc) for c in chunks])  # warm the workers
            t0 = time.perf_counter()
            pool.map(_oracle_worker, [(c, mt, mc) for c in chunks])
            dt = time.perf_counter() - t0
    return sample / dt, dt


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path (oracle port; the Python reference cannot travel
    to the GPU box) with every host core, on the same workload / metric."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(64, 8 * cores)
    total = 0
    t_all = 0.0
    for _ in range(args.warmup and 1):
        cpu_baseline(min(sample, 2 * cores), cores)
    for _ in range(args.steps):
        v, dt = cpu_baseline(sample, cores)
        total += sample
        t_all += dt
    value = total / t_all
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "macenko_transform_256", "sample_tiles_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} seeded 256x256x3 tiles per step, multiprocessing over {cores} cores"},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=8192, help="tiles per GPU per step (device-resident leg)")
    ap.add_argument("--e2e-batch", type=int, default=1024, help="tiles per GPU per step (host-buffer leg)")
    ap.add_argument("--method", default="macenko")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    ge.build()
    from tiatoolbox_b200 import _lib
    from tiatoolbox_b200.synthetic import he_target
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    norm = get_normalizer(args.method)
    norm.device = local_rank
    norm.fit(he_target(PATCH, PATCH))

    # ---- device-resident leg: B tiles per GPU (inputs + outputs >> 126 MB L2) ----
    B = args.batch
    uniq = torch.from_numpy(_tiles(128, seed=rank)).to(dev)
    x = uniq.repeat((B + 127) // 128, 1, 1, 1)[:B].contiguous()
    del uniq
    lib = _lib.lib()
    h = _lib.handle(local_rank)

    def step():
        out, info = norm.transform_batch(x, return_info=True, raise_on_error=False)
        if world > 1:  # final result gather: per-tile stain matrices + status (the tile map itself stays sharded)
            gathered = [torch.empty_like(info["stain_matrix_source"]) for _ in range(world)]
            dist.all_gather(gathered, info["stain_matrix_source"].contiguous())
        return out, info

    for _ in range(args.warmup):
        out, info = step()
    torch.cuda.synchronize()
    launches_per_step = lib.tiab_last_launch_count(h)
    assert int(info["status"].sum().item()) == 0, "synthetic tiles must all normalise"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        out, info = step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms / 1e3)

    # ---- e2e leg: numpy host buffers through the public API (pinned staging, H2D, kernels, D2H) ----
    Be = args.e2e_batch
    host_tiles = np.ascontiguousarray(x[:Be].cpu().numpy())
    for _ in range(2):
        norm.transform_batch(host_tiles, raise_on_error=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = norm.transform_batch(host_tiles, raise_on_error=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * args.steps / float(te.item())
    assert res.shape == host_tiles.shape

    if rank == 0:
        peak, peak_src = _peaks()
        med_ms = statistics.median(kernel_ms)
        achieved = B * BYTES_PER_PATCH_ALGO / (med_ms / 1e3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "stain_fused_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.method}_transform_256", "tiles_per_gpu_per_step": B, "tile": [PATCH, PATCH, 3],
                       "l2_policy": f"inputs+outputs {2 * B * PATCH * PATCH * 3 / 1e6:.0f} MB per step > 126 MB L2",
                       "parallelism": f"patch-shard x{world}"},
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": Be * PATCH * PATCH * 3,
                    "d2h_bytes_per_step": Be * PATCH * PATCH * 3 + Be * 68, "tiles_per_gpu_per_step": Be},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "stain_fused_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": B * BYTES_PER_PATCH_ALGO, "kernel_ms": med_ms},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            v, dt = cpu_baseline(64, 1)
            line["cpu_baseline"] = {"value": v, "unit": "patches/s", "cores": 1, "kind": "port",
                                    "sample": f"64 seeded 256x256x3 tiles, single process numpy oracle, {dt:.2f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
