#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the stain + patch-CNN hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload macenko_hovernet|macenko]

Metric (BASELINE.json): patches/sec on 256x256x3 tiles for Macenko stain normalisation + HoVer-Net
(hovernet_fast-pannuke architecture, seeded weights) incl. the on-device watershed post-processing.
One "step" = one pass of that path over one batch of synthetic tiles per GPU (SURVEY.md §8d generator).
At N>1 the driver launches this file under torchrun, one rank per GPU; tiles are sharded across ranks
(weak scaling: per-GPU batch fixed), no data-path collective except the all-gather of the final instance
maps.  Rank 0 prints ONE JSON line.

`value`        device-timed whole-job patches/s, inputs resident in HBM (CUDA events, max over ranks)
`e2e`          same metric through the public engine API (NucleusInstanceSegmentor.run with the stain
               normaliser installed as preproc_func) with HOST numpy buffers: H2D + kernels + D2H + host tail
`roofline`     the dominant kernel (conv_gemm_kernel, tensor-bound): algorithmic FLOPs / its summed launch
               durations (CUDA events per launch, measured live) vs MEASURED_PEAKS.json
`roofline_stain` the fused stain kernel (HBM-bound) measured the same way
`cpu_baseline` the CPU port of the reference path (numpy oracle + fp32 torch forward + post-proc oracle)
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
HOVERNET_GFLOP_PER_PATCH = 297.87                    # SURVEY.md Appendix C (2*MAC of all convs)
METRICS = {
    "macenko_hovernet": "patches/sec (256x256x3) Macenko + HoVer-Net (hovernet_fast) incl. on-device post-processing",
    "macenko": "patches/sec (256x256x3) Macenko stain-normalisation transform",
}


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


def _tiles(n_unique: int, seed: int = 0) -> np.ndarray:
    from tiatoolbox_b200.synthetic import he_tiles

    return he_tiles(n_unique, PATCH, PATCH, seed=seed)


def _hovernet_weights():
    from tiatoolbox_b200.models.architecture._params import seeded_state_dict, with_calibration
    from tiatoolbox_b200.models.architecture.hovernet import hovernet_spec

    return with_calibration(seeded_state_dict(hovernet_spec(3, 6, "fast"), seed=0), ROOT / "tests" / "golden" / "hovernet_fast_bn_calib.npz")


# ---------------------------------------------------------------------------------------------------
# CPU arm: a port of the reference's CPU path (the Python reference cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------------
def _cpu_stain_worker(args):
    from oracle import stain_oracle as so

    tiles, mt, mc = args
    return np.stack([so.transform("macenko", t.copy(), mt, mc) for t in tiles])


def cpu_pipeline(workload: str, sample: int, procs: int, threads: int):
    """Macenko (numpy oracle, `procs` processes) [+ HoVer-Net fp32 torch forward (`threads` threads) + post-proc oracle]."""
    import torch

    from oracle import stain_oracle as so
    from tiatoolbox_b200.synthetic import he_target

    tiles = _tiles(sample, seed=0)
    mt, mc = so.fit("macenko", he_target(PATCH, PATCH))
    t0 = time.perf_counter()
    if procs <= 1:
        norm = _cpu_stain_worker((tiles, mt, mc))
    else:
        import multiprocessing as mp

        chunks = [c for c in np.array_split(tiles, procs) if len(c)]
        with mp.get_context("fork").Pool(len(chunks)) as pool:
            norm = np.concatenate(pool.map(_cpu_stain_worker, [(c, mt, mc) for c in chunks]))
    t_stain = time.perf_counter() - t0
    t_net = t_post = 0.0
    if workload == "macenko_hovernet":
        from oracle import hovernet_postproc_oracle as po
        from oracle.torch_reference import hovernet_forward_reference, hovernet_heads_reference

        torch.set_num_threads(threads)
        sd = _hovernet_weights()
        t0 = time.perf_counter()
        with torch.inference_mode():
            outs = []
            for i in range(0, sample, 4):
                x = torch.from_numpy(norm[i:i + 4]).permute(0, 3, 1, 2).float()
                outs.append([o.numpy() for o in hovernet_heads_reference(hovernet_forward_reference(sd, x, "fast", 6))])
        t_net = time.perf_counter() - t0
        t0 = time.perf_counter()
        for o in outs:
            for j in range(len(o[0])):
                po.postproc(o[0][j], o[1][j], o[2][j])
        t_post = time.perf_counter() - t0
    total = t_stain + t_net + t_post
    return sample / total, {"stain_s": t_stain, "forward_s": t_net, "postproc_s": t_post}


def _best_torch_threads(cores: int) -> int:
    """oneDNN convs on these boxes slow down badly when every hyper-thread is used: give the CPU arm its best setting."""
    import torch

    from oracle.torch_reference import hovernet_forward_reference

    sd = _hovernet_weights()
    x = torch.from_numpy(_tiles(1, seed=5)).permute(0, 3, 1, 2).float()
    best, best_t = cores, float("inf")
    for th in sorted({min(cores, t) for t in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(th)
        t0 = time.perf_counter()
        with torch.inference_mode():
            hovernet_forward_reference(sd, x, "fast", 6)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = th, dt
    return best


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 8 if args.workload == "macenko_hovernet" else max(64, 8 * cores)
    threads = _best_torch_threads(cores) if args.workload == "macenko_hovernet" else cores
    total_s, n = 0.0, 0
    parts = {}
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, parts = cpu_pipeline(args.workload, sample, min(cores, sample), threads)
        if i >= args.warmup:
            total_s += time.perf_counter() - t0
            n += sample
    value = n / total_s
    line = {
        "impl": "reference", "metric": METRICS[args.workload], "value": value, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": args.workload, "sample_tiles_per_step": sample, "last_step_split_s": parts, "torch_threads": threads},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} seeded 256x256x3 tiles per step; numpy Macenko oracle over {min(cores, sample)} processes"
                                   + ("; fp32 torch HoVer-Net forward on all cores; post-proc oracle (cv2/scipy + restated skimage)"
                                      if args.workload == "macenko_hovernet" else "")},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="macenko_hovernet", choices=sorted(METRICS))
    ap.add_argument("--batch", type=int, default=None, help="tiles per GPU per step (device-resident leg)")
    ap.add_argument("--e2e-batch", type=int, default=None, help="tiles per GPU per step (host-buffer leg)")
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from tiatoolbox_b200.models.architecture.hovernet import HoVerNet
    from tiatoolbox_b200.models.engine import NucleusInstanceSegmentor
    from tiatoolbox_b200.synthetic import he_target
    from tiatoolbox_b200.tools.stainnorm import get_normalizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    combined = args.workload == "macenko_hovernet"
    B = args.batch or (16 if combined else 8192)
    # tiles per e2e call: a slide region feeds thousands of tiles per run(); 64 = four engine batches, enough for the engine's host / device
    # pipelining to reach steady state while the default bench still finishes in seconds
    Be = args.e2e_batch or (64 if combined else 1024)
    lib = _lib.lib()
    h = _lib.handle(local_rank)

    norm = get_normalizer("macenko")
    norm.device = local_rank
    norm.fit(he_target(PATCH, PATCH))
    uniq = torch.from_numpy(_tiles(min(128, max(B, 16)), seed=rank)).to(dev)
    x = uniq.repeat((B + len(uniq) - 1) // len(uniq), 1, 1, 1)[:B].contiguous()

    plan = None
    if combined:
        model = HoVerNet(num_types=6, mode="fast")
        model.load_state_dict(_hovernet_weights())
        model.precision = args.precision
        model.to(dev)
        plan = model.plan_for(B, dev, (PATCH, PATCH))

    launches = {}

    def step():
        out, info = norm.transform_batch(x, return_info=True, raise_on_error=False)
        launches["stain"] = lib.tiab_last_launch_count(h)
        if not combined:
            res = info["stain_matrix_source"]
        else:
            np_m, hv_m, tp_m = plan.forward(out)
            inst, _table = pp.proc_np_hv_batch(np_m, hv_m, tp_m)
            launches["postproc"] = lib.tiab_last_launch_count(h)
            res = inst
        if world > 1:  # the only collective: all-gather of the final result map
            gathered = torch.empty((world * res.shape[0], *res.shape[1:]), dtype=res.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, res.contiguous())
        return out, info, res

    for _ in range(args.warmup):
        out, info, res = step()
    torch.cuda.synchronize()
    assert int(info["status"].sum().item()) == 0, "synthetic tiles must all normalise"
    launches_per_step = launches.get("stain", 0) + launches.get("postproc", 0) + (plan.num_launches if plan else 0)

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
    roof = {}
    if rank == 0:
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        a, b = ev(), ev()
        ts = []
        for _ in range(5):
            a.record()
            norm.transform_batch(x, raise_on_error=False)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        roof["stain_ms"] = statistics.median(ts)
        if combined:
            conv_ms, conv_fl, conv_mma, aux_ms = 0.0, 0.0, 0.0, 0.0
            for op in plan.ops:
                a.record()
                op()
                b.record()
                torch.cuda.synchronize()
                dt = a.elapsed_time(b)
                if hasattr(op, "desc"):
                    conv_ms += dt
                    conv_fl += op.flops
                    conv_mma += op.mma_flops
                else:
                    aux_ms += dt
            maps = plan.forward(out)
            torch.cuda.synchronize()
            a.record()
            pp.proc_np_hv_batch(*maps)
            b.record()
            torch.cuda.synchronize()
            roof.update(conv_ms=conv_ms, conv_flops=conv_fl, conv_mma_flops=conv_mma, aux_ms=aux_ms, postproc_ms=a.elapsed_time(b),
                        conv_launches=sum(1 for op in plan.ops if hasattr(op, "desc")))
            # random-weight heads hold almost no nuclei, so the post-processing is ALSO timed on synthetic head maps
            # (~40 nuclei per 164-px map); reported beside the step's own figure, not part of `value`
            try:
                from tiatoolbox_b200.synthetic import hovernet_head_maps

                syn = [torch.from_numpy(m).to(dev) for m in hovernet_head_maps(B, 164, seed=0)]
                for _ in range(2):
                    pp.proc_np_hv_batch(*syn)
                torch.cuda.synchronize()
                a.record()
                for _ in range(3):
                    pp.proc_np_hv_batch(*syn)
                b.record()
                torch.cuda.synchronize()
                roof["postproc_synthetic_ms"] = a.elapsed_time(b) / 3
            except Exception as e:  # noqa: BLE001 - an auxiliary figure must never cost the bench line
                roof["postproc_synthetic_ms"] = None
                print(f"[bench] synthetic post-processing timing skipped: {e!r}", file=sys.stderr)

    # ---- e2e leg: numpy host buffers through the public API ----
    # combined workload under torchrun: every rank hands the SAME tile list to eng.run, which shards it — so the list holds Be tiles per
    # rank (weak scaling, like the device-resident leg)
    Be_total = Be * world if combined else Be
    host_tiles = np.ascontiguousarray(x[:min(Be_total, B)].cpu().numpy())
    if len(host_tiles) < Be_total:
        host_tiles = np.concatenate([host_tiles] * (-(-Be_total // len(host_tiles))))[:Be_total]
    if combined:
        eng = NucleusInstanceSegmentor(model, batch_size=B, device=str(dev))
        model.preproc_func = norm.transform
        e2e_call = lambda: eng.run(host_tiles)  # noqa: E731
        d2h = Be * 164 * 164 * 4
    else:
        e2e_call = lambda: norm.transform_batch(host_tiles, raise_on_error=False)  # noqa: E731
        d2h = world * (Be * PATCH * PATCH * 3 + Be * 68)
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
    if combined:   # bytes the engine actually copied device -> host in one call: instance maps + instance tables + contour vertices
        d2h = int(getattr(eng, "d2h_bytes", d2h))
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # total tiles per step over all ranks: eng.run shards the Be_total list itself; the stain-only call runs Be tiles on every rank
    e2e_units = Be_total if combined else world * Be
    e2e_value = e2e_units * args.steps / float(te.item())

    if rank == 0:
        pk = _peaks()
        line = {
            "metric": METRICS[args.workload], "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16x3 (split-fp16 operands, fp32 accumulate)" if args.precision == "fp16x3" else "f16",
            "data": "synthetic",
            "config": {"workload": args.workload, "tiles_per_gpu_per_step": B, "tile": [PATCH, PATCH, 3],
                       "model": "hovernet_fast (num_types=6), seeded+calibrated random weights" if combined else None,
                       "precision": args.precision if combined else "f32",
                       "l2_policy": ("activations %.1f GB per step >> 126 MB L2" % (plan.bytes_allocated() / 1e9)) if combined else
                                    f"inputs+outputs {2 * B * PATCH * PATCH * 3 / 1e6:.0f} MB per step > 126 MB L2",
                       "parallelism": f"patch-shard x{world}, all-gather of the result map"},
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": e2e_units * PATCH * PATCH * 3, "d2h_bytes_per_step": d2h,
                    "tiles_per_step": e2e_units, "api": "NucleusInstanceSegmentor.run(numpy tiles), preproc_func = MacenkoNormalizer.transform"
                    if combined else "StainNormalizer.transform_batch(numpy)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        stain_ach = B * STAIN_BYTES_PER_PATCH / (roof["stain_ms"] / 1e3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "stain_fused_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_tile")
            traffic = None if traffic is None else traffic * B
        stain_roof = {"bound": "hbm", "kernel": "stain_fused_kernel", "achieved": stain_ach, "peak": pk["hbm"], "unit": "GB/s",
                      "frac": stain_ach / pk["hbm"], "traffic": traffic, "peak_source": pk["src"],
                      "algorithmic_bytes_per_launch": B * STAIN_BYTES_PER_PATCH, "kernel_ms": roof["stain_ms"]}
        if combined:
            ach = roof["conv_flops"] / (roof["conv_ms"] / 1e3) / 1e12
            line["roofline"] = {
                "bound": "tensor", "kernel": "conv_gemm_kernel (all launches of one step)", "achieved": ach, "peak": pk["tf_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": None, "peak_source": pk["src"] + " bf16 sustained",
                "algorithmic_flops_per_step": roof["conv_flops"], "kernel_ms_per_step": roof["conv_ms"],
                "mma_issue_rate_tflops": roof["conv_mma_flops"] / (roof["conv_ms"] / 1e3) / 1e12,
                "launches_per_step": roof["conv_launches"],
                "note": "fp16x3 issues 3 tensor-core passes per algorithmic FLOP; the 1x1 layers of the network are HBM-bound, "
                        "so the aggregate sits well below the tensor peak (per-launch table: profiles/)",
            }
            line["breakdown_ms"] = {"stain": roof["stain_ms"], "conv_gemm": roof["conv_ms"], "aux_glue": roof["aux_ms"],
                                    "postproc": roof["postproc_ms"],
                                    "postproc_on_synthetic_nuclei_maps": roof.get("postproc_synthetic_ms"),
                                    "note": "per-launch CUDA events, eager (not graph) launches; the last entry times the same post-processing on "
                                            "synthetic head maps with ~40 nuclei per map (the network's random-weight heads hold almost none)"}
            line["roofline_stain"] = stain_roof
        else:
            line["roofline"] = stain_roof
        if not args.no_cpu_baseline and world == 1:
            sample = 4 if combined else 64
            v, parts = cpu_pipeline(args.workload, sample, 1, min(32, os.cpu_count() or 1))
            line["cpu_baseline"] = {"value": v, "unit": "patches/s", "cores": min(32, os.cpu_count() or 1) if combined else 1, "kind": "port",
                                    "sample": f"{sample} seeded tiles: numpy Macenko oracle (1 process)"
                                              + (", fp32 torch HoVer-Net forward (all cores), post-proc oracle" if combined else ""),
                                    "split_s": parts}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
