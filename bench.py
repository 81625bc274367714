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
    "k_fused": lambda P, U, A: 3 * P + P + P // 4,
    "k_gray_blur5": lambda P, U, A: 3 * P + P,
    "k_canny_front": lambda P, U, A: P + P // 4,
    "k_adaptive": lambda P, U, A: P + P,
    "k_color_bitmap": lambda P, U, A: 3 * P,
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
    ap.add_argument("--workload", choices=sorted(WORKLOAD_BYTES), default="pipeline")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--k", type=int, default=5, help="palette size (n_colors); BASELINE config 3 uses 16")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic images generated on the host")
    ap.add_argument("--e2e-steps", type=int, default=2)
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


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.cpu_seconds)     # before CUDA is initialised (fork-based pool)

    import numpy as np
    import torch
    import torch.distributed as dist

    from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
    from low_level_feature_extraction_b200.synth import design_image

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, H, W = args.batch, args.height, args.width
    P = H * W

    # synthetic inputs: a few distinct design images per rank, tiled (rolled) to the batch on the device
    base = np.stack([design_image(H, W, 100 * rank + s) for s in range(args.distinct)])
    base_d = torch.from_numpy(base).to(dev)
    batch = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    for i in range(B):
        batch[i] = torch.roll(base_d[i % args.distinct], shifts=7 * (i // args.distinct), dims=0)
    del base_d

    cfg = BatchConfig(colors=args.workload in ("pipeline", "colors", "palette_shadows"),
                      shapes=args.workload in ("pipeline", "shapes"),
                      shadows=args.workload in ("pipeline", "shadows", "palette_shadows"), k=args.k)
    an = BatchAnalyzer(local, H, W, cfg)
    eng = an.engines[0]
    out = an.alloc_outputs(B)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None   # started early: samples during the warm-up are the fallback
    for _ in range(args.warmup):
        an.run_device(batch, out)
    barrier()
    if sampler:
        sampler.wait_first()
    barrier()
    launches0 = eng.launches
    eng.ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        an.run_device(batch, out)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    kernels = eng.ctx.profile_end()
    launches = eng.launches - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- end to end through the host-buffer API ----------------------------------------
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        host_in.copy_(batch)
        host_out = an.alloc_host_outputs(B)
        an.run_host(host_in, host_out)     # warm-up (allocates the staging chunks)
        barrier()
        te = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = an.run_host(host_in, host_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - te
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": B * world * args.e2e_steps / dt, "unit": "images/sec",
               "h2d_bytes_per_step": int(res["_h2d_bytes"]) * world, "d2h_bytes_per_step": int(res["_d2h_bytes"]) * world,
               "steps": args.e2e_steps,
               "api": "BatchAnalyzer.run_host: pinned host images in, host masks/palettes out, chunked copies "
                      f"overlapped with kernels ({an.cfg.host_streams} streams, {an.cfg.host_chunk} images per stage)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        u_avg = float(out["count"].float().mean().item()) if "count" in out else 0.0

        def ncu_traffic(name, imgs_per_launch):
            """dram bytes per launch from the committed ncu capture, scaled to this run's images per launch."""
            rec = NCU_TRAFFIC.get(name)
            if not rec:
                return None
            scale = imgs_per_launch / rec["images_per_launch"] if rec.get("images_per_launch") else 1.0
            return rec["bytes_per_launch"] * scale

        def kernel_bytes(name):
            return KERNEL_BYTES.get(name, lambda P, U, A: 0)(P, u_avg, cfg.attempts)

        # dominant kernel: largest total device time inside the timed region
        dom = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else (None, None)
        roofline = None
        total_kernel_ms = sum(v["ms"] for v in kernels.values()) or 1.0
        if dom[0]:
            name, rec = dom
            per_launch_ms = rec["ms"] / rec["launches"]
            imgs_per_launch = B * args.steps / rec["launches"]
            algo = kernel_bytes(name) * imgs_per_launch
            ach = algo / (per_launch_ms / 1e3) / 1e9
            roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": ncu_traffic(name, imgs_per_launch), "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
                        "avg_launch_ms": per_launch_ms, "share_of_step": rec["ms"] / total_kernel_ms,
                        "images_per_launch": imgs_per_launch}
        step_bytes = WORKLOAD_BYTES[args.workload](P) * B
        step_ach = step_bytes / (ms / args.steps / 1e3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_of(args, world),
                "roofline": roofline,
                "roofline_step": {"bound": "hbm", "algorithmic_bytes_per_image": WORKLOAD_BYTES[args.workload](P),
                                  "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                                  "note": "whole step (all kernels) against the workload's algorithmic bytes, per GPU"},
                "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                                "achieved_gbs": kernel_bytes(k) * B * args.steps / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else 0.0}
                            for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
                "unique_colours_per_image": u_avg,
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
