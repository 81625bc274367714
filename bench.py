#!/usr/bin/env python
"""Headline benchmark: images/sec @1080p for the image hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pipeline|shapes|shadows|colors]
    python bench.py --impl reference ...      # the reference's own CPU path on this box's host cores

A step is one pass of the hot path over one batch of synthetic design-style images per
GPU (weak scaling: the per-GPU batch is fixed; batches shard by image with no collective).
`value` is device-resident throughput (inputs already in HBM), `e2e` is the same metric
through the host-buffer API with the host<->device copies inside the timed region.
One JSON line on stdout (rank 0).  See DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec @1080p (colors+edges)"

# Algorithmic (compulsory) HBM bytes per image per kernel launch: what the kernel must read once +
# write once (P = H*W pixels, U = unique colours of the image, A = k-means attempts).  DESIGN.md
# "Kernels" derives each figure.
BM = 1 << 21                                   # the 2^24-bit colour bitmap of one image
KERNEL_BYTES = {
    # fused front: read BGR (3P); write shadow mask (P) + weak/strong bit planes (2 * P/8); the
    # bitmap is L2-resident scratch (zeroed + scanned by other kernels) and is not counted here
    # front kernel: read BGR (3P); write the weak/strong bit planes (2 * P/8) + the blurred gray plane (P)
    "k_fused": lambda P, U, A: 3 * P + P + P // 4,
    # shadow kernel: read the blurred plane, write the mask
    "k_shadow": lambda P, U, A: P + P,
    "k_gray_blur5": lambda P, U, A: 3 * P + P,
    "k_canny_front": lambda P, U, A: P + P // 4,
    "k_adaptive": lambda P, U, A: P + P,
    "k_color_pass": lambda P, U, A: 3 * P,
    # hysteresis: read weak + strong planes, write the edge plane (in place)
    "k_hyst": lambda P, U, A: 3 * (P // 8),
    "k_hyst_mask": lambda P, U, A: 2 * (P // 8) + P,
    "k_hyst_strips": lambda P, U, A: 3 * (P // 8),
    "k_hyst_strips_again": lambda P, U, A: 0,   # revisits flagged strips only: no compulsory traffic
    "k_hyst_finish": lambda P, U, A: 0,
    "k_plane_to_mask_dilate": lambda P, U, A: P // 8 + P,
    "k_plane_to_mask": lambda P, U, A: P // 8 + P,
    "k_bm_blocksum": lambda P, U, A: BM,
    "k_bm_blockscan": lambda P, U, A: 4096,
    "k_bm_emit": lambda P, U, A: BM + 4 * U,
    # k-means: every (attempt, image) CTA reads the key list once and writes its labels once;
    # all iterations run out of shared memory, so this kernel is SM-bound, not HBM-bound
    "k_kmeans_fast": lambda P, U, A: A * (4 * U + U),
    "k_kmeans_fast_long": lambda P, U, A: 0,       # same images counted under k_kmeans_fast
    "k_kmeans_fast_global": lambda P, U, A: 0,
    "k_kmeans_pp": lambda P, U, A: A * 4 * U,
    "k_kmeans_lloyd": lambda P, U, A: A * (4 * U + U),
    "k_kmeans": lambda P, U, A: A * (4 * U + U),
    "k_kmeans_pick": lambda P, U, A: A * U + 4 * U,
}

# Algorithmic bytes per image of a whole step: read the BGR input once,
# write each required output once.
WORKLOAD_BYTES = {
    "pipeline": lambda P: 3 * P + 2 * P,   # edge mask + shadow mask           (config 4: 5P)
    "shapes": lambda P: 3 * P + P,         # dilated edge mask                 (config 2: 4P)
    "shadows": lambda P: 3 * P + P,
    "colors": lambda P: 3 * P,
    "palette_shadows": lambda P: 3 * P + P,   # BASELINE config 3: palette + shadow threshold mask
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOAD_BYTES) + ["pixel_kmeans"], default="pipeline")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--k", type=int, default=5, help="palette size (n_colors); BASELINE config 3 uses 16")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic images generated on the host")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-tail", choices=["thread", "inline"], default="thread",
                    help="host palette tail of e2e: on a worker thread under the next step's copies, or inline")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra records (BASELINE configs 1, 2, 3, literal 4, 5, contours, adversarial frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def config_of(args, world):
    return {
        "workload": f"{args.workload}: batch of {args.batch} {args.width}x{args.height} synthetic design images per GPU"
                    + {"pipeline": " through colours (k-means k=5 over unique colours, 10 kmeans++ attempts) + shape mask"
                                   " (gray/blur/Canny/dilate) + shadow mask (gray/blur/adaptive threshold)",
                       "shapes": " through preprocess + Canny edge/shape masks (BASELINE config 2)",
                       "shadows": " through the shadow threshold mask",
                       "colors": f" through the colour palette (k={args.k})",
                       "palette_shadows": f" through the colour palette (k={args.k}) + shadow threshold mask (BASELINE config 3)"}[args.workload]
                    .replace("k-means k=5", f"k-means k={args.k}"),
        "images_per_gpu": args.batch, "height": args.height, "width": args.width,
        "global_images_per_step": args.batch * world,
        "parallelism": f"dp{world} (images sharded by rank, no collective)",
        "l2": "inputs larger than L2 (batch is %.0f MB per GPU vs 126 MB L2); no explicit flush" %
              (args.batch * args.height * args.width * 3 / 1e6),
    }


# ---------------------------------------------------------------------------------------
def cpu_baseline(args, seconds):
    from oracle.refbench import CpuReference

    ref = CpuReference(args.workload, args.height, args.width, k=args.k)
    n = ref.images_per_step(seconds)
    dt = ref.step(n)
    ref.close()
    return {"value": n / dt, "unit": "images/sec", "cores": ref.procs, "kind": ref.kind,
            "sample": f"{n} images of {args.width}x{args.height} ({args.workload}) in {dt:.2f} s; one image per task over "
                      f"{ref.procs} processes, cv2.setNumThreads(1) each"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle.refbench import CpuReference

    ref = CpuReference(args.workload, args.height, args.width, k=args.k)
    n = ref.images_per_step(max(2.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        ref.step(n)
    t = [ref.step(n) for _ in range(args.steps)]
    ref.close()
    total = sum(t)
    value = n * args.steps / total
    line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference", "config": config_of(args, world),
            "cpu_baseline": {"value": value, "unit": "images/sec", "cores": ref.procs, "kind": ref.kind,
                             "sample": f"{n} images per step, one image per task over {ref.procs} processes"},
            "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs a moment before its first line: do not start the timed region without a live sampler."""
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end and self.proc.poll() is None:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# (profiles/ncu_r1_summary.md), bytes; kernels not captured yet report null.
NCU_TRAFFIC = {}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _f:
        NCU_TRAFFIC = json.load(_f)
except Exception:
    pass


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- measurement helpers ------------------------------------------------------------------------------
# kernels whose time is not bounded by HBM: reported with the bound that does limit them
SM_BOUND = {"k_kmeans_fast": "sm (issue-bound: every Lloyd iteration runs out of shared memory, zero HBM traffic per iteration)",
            "k_kmeans_fast_long": "sm", "k_kmeans_fast_global": "sm / L2 (lists that do not fit in shared memory)",
            "k_hyst_mask": "latency (data-dependent propagation inside a thread-block cluster)"}


def device_batch(dev, rank, B, H, W, distinct, kind="design"):
    """(B, H, W, 3) u8 device batch: `distinct` synthetic images generated on the host, tiled (row-rolled) to B."""
    import numpy as np
    import torch

    from low_level_feature_extraction_b200.synth import design_image, noise_image

    distinct = max(1, min(distinct, B))
    gen = design_image if kind == "design" else noise_image
    batch = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    for s in range(distinct):
        batch[s] = torch.from_numpy(gen(H, W, 100 * rank + s)).to(dev)
    for i in range(distinct, B):
        batch[i] = torch.roll(batch[i % distinct], shifts=7 * (i // distinct), dims=0)
    return batch


def timed_steps(an, batch, out, steps, warmup, world, dev, profile=False):
    """-> (ms per step [max over ranks], per-kernel dict or None, launches summed over ranks)."""
    import torch
    import torch.distributed as dist

    eng = an.engines[0]
    for _ in range(warmup):
        an.run_device(batch, out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launches
    if profile:
        eng.ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        an.run_device(batch, out)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kernels = eng.ctx.profile_end() if profile else None
    ms = e0.elapsed_time(e1) / steps
    launches = eng.launches - launches0
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches = float(mx[0].item()), int(t[1].item())
    return ms, kernels, launches


def measure_e2e(an, batch, steps, world, dev, colors, tail="thread"):
    """images/sec through BatchAnalyzer.run_host: pinned host images in, host masks + palettes out, the
    reference's palette tail (argsort / hex / ColorFeatures, color_extractor.py:231-284) computed on the host for
    every image -- on a worker thread, while the next step's copies and kernels run."""
    import time
    from concurrent.futures import ThreadPoolExecutor

    import torch
    import torch.distributed as dist

    B = batch.shape[0]
    host_in = torch.empty(tuple(batch.shape), dtype=torch.uint8).pin_memory()
    host_in.copy_(batch)
    outs = [an.alloc_host_outputs(B) for _ in range(2)]      # double-buffered: the tail of step i reads buffer i % 2
    res = an.run_host(host_in, outs[0])                      # warm-up (allocates the staging chunks)
    an.run_host_async(host_in, outs[1]).result()             # ... and the pinned bit planes of the second batch in flight
    if colors:
        an.palettes(res)                                     # ... and the tail's first call (imports, pydantic model build)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    pool = ThreadPoolExecutor(1)
    sys.setswitchinterval(0.0005)   # the tail thread must not hold the interpreter for 5 ms while this thread enqueues copies
    # The region is wall-clock on the host (copies, host threads and the interpreter are part of what is measured), so
    # one scheduling hiccup of the box moves a 0.2 s block by tens of per cent: three blocks of `steps` steps are
    # timed, the MEDIAN block is the value and all three are in the record.
    import gc

    gc.collect()
    gc.disable()     # a generation-2 pass over the interpreter's heap in the middle of a 0.2 s block is not part of the workload
    blocks = []
    for _ in range(3):
        pending = []
        n_palettes = 0
        te = time.perf_counter()
        prev = None

        def finish(call):
            nonlocal n_palettes
            res_ = call.result()                                 # this step's masks / palettes are on the host
            if colors and tail == "thread":
                pending.append(pool.submit(an.palettes, res_))
            elif colors:
                n_palettes += len(an.palettes(res_))
            return res_

        for i in range(steps):
            while pending:                                       # the tail of step i - 2 still reads buffer i % 2
                n_palettes += len(pending.pop(0).result())
            call = an.run_host_async(host_in, outs[i % 2])       # step i is enqueued while step i - 1 drains
            if prev is not None:
                res = finish(prev)
            prev = call
        res = finish(prev)
        for f in pending:
            n_palettes += len(f.result())
        torch.cuda.synchronize()
        dt = time.perf_counter() - te
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        blocks.append(dt)
    pool.shutdown()
    gc.enable()
    dt = sorted(blocks)[1]
    return {"value": B * world * steps / dt, "unit": "images/sec",
            "h2d_bytes_per_step": int(res["_h2d_bytes"]) * world, "d2h_bytes_per_step": int(res["_d2h_bytes"]) * world,
            "steps": steps, "blocks_images_per_sec": [round(B * world * steps / b, 1) for b in blocks],
            "note": "median of three blocks of `steps` batches; Python's cyclic garbage collector is paused during the blocks",
            "palette_tail_on_host": (tail if colors else False), "palettes_per_step": n_palettes // max(1, steps),
            "api": "BatchAnalyzer.run_host_async (two batches in flight): pinned host images in, host masks + ColorFeatures out; "
                   f"chunked copies overlapped with kernels ({an.cfg.host_streams} streams, {an.cfg.host_chunk} images per stage)"}


def pipeline_record(args, workload, B, H, W, k, steps, warmup, rank, world, local, dev, distinct, kind="design", e2e_steps=0,
                    max_unique=1 << 16, sampler=None):
    """One workload through BatchAnalyzer: device-resident throughput, per-kernel times (a separate short pass with the
    two chains serialised, so that a kernel's time is its own), roofline of the dominant HBM-path kernel and of the step."""
    import torch

    from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig

    P = H * W
    batch = device_batch(dev, rank, B, H, W, distinct, kind)
    cfg = BatchConfig(colors=workload in ("pipeline", "colors", "palette_shadows"),
                      shapes=workload in ("pipeline", "shapes"),
                      shadows=workload in ("pipeline", "shadows", "palette_shadows"), k=k, max_unique=max_unique)
    an = BatchAnalyzer(local, H, W, cfg)
    eng = an.engines[0]
    out = an.alloc_outputs(B)
    t_begin = time.perf_counter()
    ms, _, launches = timed_steps(an, batch, out, steps, warmup, world, dev)
    t_end = time.perf_counter()
    # the nvidia-smi sampler covers the warm-up + timed steps only: its NVML queries (every 50 ms) stall CUDA submissions
    # enough to cost the copy-bound e2e loop two thirds of its throughput
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    # per-kernel times: serial pass (on two streams the kernels of the two chains overlap and each one's event time
    # contains the other's share of the SMs)
    eng.ctx.set_option("serial", 1)
    ms_serial, kernels, _ = timed_steps(an, batch, out, max(2, min(steps, 5)), 1, world, dev, profile=True)
    eng.ctx.set_option("serial", 0)
    psteps = max(2, min(steps, 5))
    peak, peak_src = measured_peak()
    u_avg = float(out["count"].float().mean().item()) if "count" in out else 0.0
    overflow = int((out["count"] > cfg.max_unique).sum().item()) if "count" in out else 0

    def kbytes(name):
        return KERNEL_BYTES.get(name, lambda P, U, A: 0)(P, u_avg, cfg.attempts)

    krec = {}
    for name, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]):
        per_step = v["ms"] / psteps
        gbs = kbytes(name) * B / (per_step / 1e3) / 1e9 if per_step > 0 else 0.0
        krec[name] = {"ms_per_step": per_step, "launches_per_step": v["launches"] / psteps,
                      "bound": SM_BOUND.get(name, "hbm"), "achieved_gbs": gbs,
                      "frac_of_hbm_peak": gbs / peak}
    hbm = [(n, r) for n, r in krec.items() if r["bound"] == "hbm" and kbytes(n) > 0]
    roofline = None
    if hbm:
        name, r = max(hbm, key=lambda kv: kv[1]["ms_per_step"])
        per_launch_ms = r["ms_per_step"] / r["launches_per_step"]
        imgs = B / r["launches_per_step"]
        algo = kbytes(name) * imgs
        ach = algo / (per_launch_ms / 1e3) / 1e9
        rec = NCU_TRAFFIC.get(name)
        traffic = None
        if rec:
            traffic = rec["bytes_per_launch"] * (imgs / rec["images_per_launch"] if rec.get("images_per_launch") else 1.0)
        total = sum(x["ms_per_step"] for x in krec.values()) or 1.0
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
                    "avg_launch_ms": per_launch_ms, "share_of_step": r["ms_per_step"] / total, "images_per_launch": imgs,
                    "note": "dominant kernel of the HBM path (largest CUDA-event time among the kernels that stream the images); "
                            "per-kernel times from a pass with the two chains on one stream"}
    step_bytes = WORKLOAD_BYTES[workload](P)
    step_ach = step_bytes * B / (ms / 1e3) / 1e9
    rec = {"workload": workload, "images_per_gpu": B, "height": H, "width": W, "k": k, "input": kind,
           "distinct_images": min(distinct, B), "value": B * world / (ms / 1e3), "unit": "images/sec", "ms_per_step": ms,
           "ms_per_step_serial": ms_serial, "steps": steps, "gpu_launches": launches, "unique_colours_per_image": u_avg,
           "lists_redone_per_step": overflow, "clocks": clocks,
           "roofline": roofline,
           "roofline_step": {"bound": "hbm", "algorithmic_bytes_per_image": step_bytes, "achieved": step_ach, "peak": peak,
                             "unit": "GB/s", "frac": step_ach / peak,
                             "note": "whole step (all kernels, both streams) against the workload's algorithmic bytes, per GPU"},
           "kernels": krec}
    if e2e_steps:
        rec["e2e"] = measure_e2e(an, batch, e2e_steps, world, dev, cfg.colors, getattr(args, "e2e_tail", "thread"))
    del batch, out, an
    torch.cuda.empty_cache()
    return rec


def contours_record(local, dev, rank, B, H, W):
    """external contours (cv2.findContours(EXTERNAL, SIMPLE) + the reference's area filter) of the shape
    masks of a device-resident batch, all images traced concurrently; next to cv2 on one host core."""
    import cv2
    import torch

    import low_level_feature_extraction_b200 as pkg

    eng = pkg.engine(local)
    batch = device_batch(dev, rank, B, H, W, 32, "design")
    masks = eng.shape_mask(batch)
    hdr, pts, cnt = eng.contours_external(masks, 200, 1024, 8192)      # warm-up; the timed calls reuse these buffers
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        eng.ctx.call("llfe_contours_external", masks, B, H, W, 200, hdr, 1024, pts, 8192, cnt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    c = cnt.cpu()
    host = masks[:8].cpu().numpy()
    t0 = time.perf_counter()
    for m in host:
        cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cv_ms = (time.perf_counter() - t0) * 1e3 / len(host)
    rec = {"workload": f"external contours of {B} {W}x{H} shape masks (device-resident), area filter 100", "images_per_gpu": B,
           "ms_per_step": ms, "value": B / (ms / 1e3), "unit": "images/sec",
           "contours_per_image": float(c[:, 0].float().mean()), "points_per_image": float(c[:, 1].float().mean()),
           "overflowed_images": int(((c[:, 0] > 1024) | (c[:, 2] != 0)).sum()),
           "cv2_findContours_ms_per_image_one_core": cv_ms,
           "d2h_bytes_per_image_if_fetched": 1024 * 40 + 8192 * 8 + 16}
    del batch, masks, hdr, pts
    torch.cuda.empty_cache()
    return rec


def config1_record():
    """BASELINE config 1, the reference's own CPU-runnable case: ONE 1080p PNG through the drop-in service calls
    (validate_and_preprocess_image = PNG decode + resize policy, extract_colors, analyze_shapes, analyze_shadow_level),
    single-image latency next to the reference's call sequence on the host cores (the port, like `cpu_baseline`).
    Runs `tools/service_latency.py` in a child process: the services keep a library context of their own."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "service_latency.py"), "--reps", "5"],
                       capture_output=True, text=True, timeout=150, cwd=ROOT)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.strip()[-300:])
    d = json.loads(r.stdout.strip().splitlines()[-1])
    return {"workload": "config 1: " + d["image"] + ", one image at a time through the drop-in services", "unit": d["unit"],
            "ours_gpu": d["ours_gpu"], "reference_cpu_port": d["reference_cpu_port"], "host_threads_cv2": d["host_threads_cv2"],
            "images_per_sec_decode_plus_palette": d["config1_images_per_sec"]}


def pixel_kmeans_record(rank, world, local, dev, height, width, k, synth="design"):
    """BASELINE config 5: ONE height x width image, rows sharded over the ranks (strong scaling), per-pixel k-means with
    exact integer sums; the colour count table crosses NVLink once (reduce-scatter), the K x 4 sums every iteration."""
    import hashlib
    import importlib.util

    import torch
    import torch.distributed as dist

    import low_level_feature_extraction_b200 as pkg
    from low_level_feature_extraction_b200.dist import PixelKMeans, row_shard

    spec = importlib.util.spec_from_file_location("run_pixel_kmeans", os.path.join(ROOT, "tools", "run_pixel_kmeans.py"))
    rp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rp)
    eng = pkg.engine(local)
    r0, r1 = row_shard(height, rank, world)
    gen = rp.synth_rows_photo if synth == "photo" else rp.synth_rows
    rows = torch.cat([gen(s, min(r1, s + 1024), width, dev) for s in range(r0, r1, 1024)], dim=0)
    g = torch.Generator().manual_seed(42)
    pos = torch.randint(0, height * width, (k,), generator=g)
    init = torch.zeros((k, 3), dtype=torch.float32, device=dev)
    for j, p in enumerate(pos.tolist()):
        y, x = divmod(p, width)
        if r0 <= y < r1:
            init[j] = rows[y - r0, x].flip(0).to(torch.float32)
    if world > 1:
        dist.all_reduce(init)
    km = PixelKMeans(eng)
    km.fit(rows, init, index_base=r0 * width)          # warm-up (allocations, NCCL channels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        res = km.fit(rows, init, index_base=r0 * width)
    e1.record()
    torch.cuda.synchronize()
    fit_ms = e0.elapsed_time(e1) / reps
    # the same fit with ncclAllReduce + update per iteration instead of the fused NVLink exchange (for comparison)
    fit_nccl_ms = None
    if world > 1:
        km2 = PixelKMeans(eng, p2p=False)
        km2.fit(rows, init, index_base=r0 * width)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(reps):
            res2 = km2.fit(rows, init, index_base=r0 * width)
        e1.record()
        torch.cuda.synchronize()
        fit_nccl_ms = e0.elapsed_time(e1) / reps
        assert torch.equal(res2.centers, res.centers) and res2.iters == res.iters
    # the all-reduce of the K x 4 sums alone (what every iteration would pay on top of its kernels with NCCL)
    sums = torch.zeros((k, 4), dtype=torch.int64, device=dev)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        km._allreduce(sums, dist.ReduceOp.SUM)
    a0.record()
    for _ in range(20):
        km._allreduce(sums, dist.ReduceOp.SUM)
    a1.record()
    torch.cuda.synchronize()
    ar_ms = a0.elapsed_time(a1) / 20 if world > 1 else 0.0
    if world > 1:
        t = torch.tensor([fit_ms, ar_ms, fit_nccl_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fit_ms, ar_ms, fit_nccl_ms = t.tolist()
    peak, _ = measured_peak()
    npix = height * width
    rec = {"workload": f"per-pixel k-means of ONE {width}x{height} image ({synth}-like synthetic rows), K={k}, rows sharded over "
                       f"{world} GPU(s); colour-histogram form (rows read once, Lloyd over the distinct colours)",
           "scaling": "strong", "n_gpus": world, "iterations": res.iters, "fit_ms": fit_ms, "value": 1e3 / fit_ms,
           "unit": "fits/sec", "iterations_per_sec": res.iters / (fit_ms / 1e3),
           "sums_exchange": ("fused: one kernel per rank stores the K x 4 u64 partial sums into every peer's mailbox over NVLink "
                             "(CUDA IPC), waits for the peers' flags, adds them and updates the centres "
                             "(llfe_kmeans_update_p2p)") if world > 1 else "none (1 GPU)",
           "fit_ms_with_nccl_allreduce": fit_nccl_ms,
           "nccl_allreduce_ms_per_iteration": ar_ms,
           "exchange_saving_per_iteration_ms": ((fit_nccl_ms - fit_ms) / max(1, res.iters)) if fit_nccl_ms else 0.0,
           "table_exchange": "reduce-scatter of the block-transposed 64 MiB colour table (each rank receives its 1/G)"
                             if world > 1 else "none (1 GPU)",
           "roofline_fit": {"bound": "hbm", "algorithmic_bytes": 3 * npix, "achieved": 3 * npix / (fit_ms / 1e3) / 1e9 / 1.0,
                            "peak": peak * world, "unit": "GB/s", "frac": 3 * npix / (fit_ms / 1e3) / 1e9 / (peak * world),
                            "note": "the whole fit against ONE read of the image (3 bytes per pixel), all GPUs"},
           "centres_sha256": hashlib.sha256(res.centers.cpu().numpy().tobytes()).hexdigest(),
           "comm_nranks_ok": (dist.get_world_size() == world) if world > 1 else True}
    del rows
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload != "pixel_kmeans":
        cpu = cpu_baseline(args, args.cpu_seconds)     # before CUDA is initialised (fork-based pool)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    if args.workload == "pixel_kmeans":
        rec = pixel_kmeans_record(rank, world, local, dev, args.height if args.height != 1080 else 16384,
                                  args.width if args.width != 1920 else 16384, args.k if args.k != 5 else 16)
        if rank == 0:
            line = {"metric": "fits/sec of one 16384x16384 per-pixel k-means (BASELINE config 5)", "value": rec["value"],
                    "unit": "fits/sec", "n_gpus": world, "steps": 5, "warmup": 1, "ms_per_step": rec["fit_ms"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                    "config": {"workload": rec["workload"]}, "record": rec}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    sampler = ClockSampler(local) if rank == 0 else None   # samples during the whole run; the main record's window is cut out
    if sampler:
        sampler.wait_first()
    main = pipeline_record(args, args.workload, args.batch, args.height, args.width, args.k, args.steps, args.warmup, rank,
                           world, local, dev, args.distinct, e2e_steps=0 if args.no_e2e else args.e2e_steps, sampler=sampler)
    clocks = main.pop("clocks")

    extra = {}
    if args.workload == "pipeline" and not args.no_extras:
        def guarded(name, fn):
            try:
                extra[name] = fn()
                extra[name].pop("clocks", None)
            except Exception as e:   # an extra record never takes the headline line down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                if world > 1:
                    raise

        guarded("config2_shapes_256x1080p", lambda: pipeline_record(
            args, "shapes", 256, 1080, 1920, 5, 5, 2, rank, world, local, dev, 32))
        # BASELINE config 4 literally: 8192 images over 8 GPUs = 1024 per GPU (the headline line runs 256 per GPU)
        guarded("config4_pipeline_1024_per_gpu_1080p", lambda: pipeline_record(
            args, "pipeline", 1024, 1080, 1920, 5, 3, 1, rank, world, local, dev, 32))
        guarded("config3_palette_shadows_64x4k_k16", lambda: pipeline_record(
            args, "palette_shadows", 64, 2160, 3840, 16, 3, 1, rank, world, local, dev, 8))
        if world == 1:
            guarded("adversarial_uniform_noise_1080p", lambda: pipeline_record(
                args, "pipeline", 4, 1080, 1920, 5, 1, 1, rank, world, local, dev, 4, kind="noise"))
        if world == 1:
            guarded("contours_256x1080p", lambda: contours_record(local, dev, rank, 256, 1080, 1920))
            guarded("config1_single_image_services_1080p", config1_record)
        guarded("config5_pixel_kmeans_16384x16384_k16", lambda: pixel_kmeans_record(
            rank, world, local, dev, 16384, 16384, 16))
        guarded("config5_pixel_kmeans_photo_like", lambda: pixel_kmeans_record(
            rank, world, local, dev, 16384, 16384, 16, synth="photo"))

    if rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": "images/sec", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_of(args, world),
                "roofline": main["roofline"], "roofline_step": main["roofline_step"], "kernels": main["kernels"],
                "ms_per_step_serial": main["ms_per_step_serial"],
                "unique_colours_per_image": main["unique_colours_per_image"],
                "cpu_baseline": cpu, "e2e": main.get("e2e"), "gpu_launches": main["gpu_launches"],
                "clocks": clocks, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
