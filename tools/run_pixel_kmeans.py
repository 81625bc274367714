#!/usr/bin/env python
"""BASELINE config 5: per-pixel k-means (K = 16) of ONE oversized image whose rows are sharded across
the ranks, with an all-reduce of the K x 4 exact sums per Lloyd iteration (NCCL over NVLink).

    python tools/run_pixel_kmeans.py [--height 16384 --width 16384 --k 16]            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/run_pixel_kmeans.py ...                                                # N GPUs

Every rank synthesises its own rows (a function of the global row index, so the image is the same
for every world size), runs `dist.PixelKMeans`, and rank 0 prints one JSON line: iterations, time per
iteration (CUDA events, max over ranks), the share of the all-reduce, the HBM roofline fraction of the
assignment kernel (3 bytes per pixel per iteration), and a checksum of the final centres that must be
identical for every world size (exact integer sums => bit-identical centres).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def synth_rows(r0: int, r1: int, w: int, device) -> torch.Tensor:
    """Design-like synthetic rows: smooth colour fields + blocks + noise, a pure function of (row, col)."""
    y = torch.arange(r0, r1, device=device, dtype=torch.int64).view(-1, 1)
    x = torch.arange(w, device=device, dtype=torch.int64).view(1, -1)
    h = (y * 1315423911 + x * 2654435761) & 0xFFFFFFFF
    h = (h ^ (h >> 15)) * 2246822519 & 0xFFFFFFFF
    noise = ((h >> 13) & 7) - 3
    blk = ((y // 512) * 131 + (x // 768) * 71) % 11
    b = (40 + blk * 19 + (x // 64) % 5 + noise).clamp(0, 255)
    g = (230 - blk * 17 + (y // 96) % 7 + noise).clamp(0, 255)
    r = (128 + ((blk * 37) % 97) + noise).clamp(0, 255)
    return torch.stack([b, g, r], dim=-1).to(torch.uint8).contiguous()


def synth_rows_photo(r0: int, r1: int, w: int, device) -> torch.Tensor:
    """Photo-like rows: smooth, differently-phased colour ramps + 5 bits of per-channel noise -- millions of
    distinct colours (the hard case for the colour-histogram form), a pure function of (row, col)."""
    y = torch.arange(r0, r1, device=device, dtype=torch.int64).view(-1, 1)
    x = torch.arange(w, device=device, dtype=torch.int64).view(1, -1)
    h = (y * 1315423911 + x * 2654435761) & 0xFFFFFFFF
    h = (h ^ (h >> 15)) * 2246822519 & 0xFFFFFFFF
    h = h ^ (h >> 13)
    b = ((x * 3 + y) // 197 + (h & 31)) % 256
    g = ((x + y * 5) // 311 + ((h >> 5) & 31)) % 256
    r = ((x * 7 + y * 2) // 523 + ((h >> 10) & 31)) % 256
    return torch.stack([b, g, r], dim=-1).to(torch.uint8).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synth", default="design", choices=["design", "photo"],
                    help="design: flat blocks + small noise (few thousand colours); photo: ramps + noise (millions)")
    ap.add_argument("--height", type=int, default=16384)
    ap.add_argument("--width", type=int, default=16384)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--max-iter", type=int, default=200)
    ap.add_argument("--per-pixel", action="store_true",
                    help="re-read the rows every iteration (k_pixels_step) instead of the colour-histogram form")
    ap.add_argument("--labels", action="store_true", help="also produce the per-pixel labels inside the timed fit")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import low_level_feature_extraction_b200 as pkg
    from low_level_feature_extraction_b200.dist import PixelKMeans, row_shard

    eng = pkg.engine(local)
    r0, r1 = row_shard(a.height, rank, world)
    synth = synth_rows_photo if a.synth == "photo" else synth_rows
    rows = torch.cat([synth(s, min(r1, s + 1024), a.width, dev) for s in range(r0, r1, 1024)], dim=0)
    # seeded initial centroids: K pixels at fixed global positions (identical on every rank)
    g = torch.Generator().manual_seed(42)
    pos = torch.randint(0, a.height * a.width, (a.k,), generator=g)
    init = torch.zeros((a.k, 3), dtype=torch.float32, device=dev)
    for j, p in enumerate(pos.tolist()):
        y, x = divmod(p, a.width)
        if r0 <= y < r1:
            init[j] = rows[y - r0, x].flip(0).to(torch.float32)
    if world > 1:
        dist.all_reduce(init)

    km = PixelKMeans(eng, histogram=not a.per_pixel)
    # time the two device parts of an iteration separately (CUDA events), then the whole fit
    sums = torch.zeros((a.k, 4), dtype=torch.int64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(2):
        sums.zero_()
        eng.kmeans_pixels_step(rows, init, sums)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        eng.kmeans_pixels_step(rows, init, sums)
    ev[1].record()
    for _ in range(20):
        km._allreduce(sums, dist.ReduceOp.SUM)
    ev[2].record()
    torch.cuda.synchronize()
    step_ms = ev[0].elapsed_time(ev[1]) / 5
    ar_ms = ev[1].elapsed_time(ev[2]) / 20

    # the one-off parts of the histogram form: count table over the rows, compaction, one step over the colours
    hist_ms = compact_ms = hstep_ms = 0.0
    n_colours = 0
    if not a.per_pixel:
        hist = torch.zeros((1 << 24,), dtype=torch.int32, device=dev)
        eng.pixels_histogram(rows, hist)
        hist.zero_()
        torch.cuda.synchronize()
        ev[0].record()
        eng.pixels_histogram(rows, hist)
        ev[1].record()
        km._allreduce(hist, dist.ReduceOp.SUM)
        keys, counts = eng.histogram_compact(hist, rank, world)  # host-sized lists for the stand-alone timing
        ev[2].record()
        for _ in range(10):
            eng.kmeans_hist_step(keys, counts, init, sums)
        ev[3].record()
        torch.cuda.synchronize()
        hist_ms, compact_ms, hstep_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]) / 10
        n_colours = keys.numel()
        del hist, keys, counts

    km.fit(rows, init, index_base=r0 * a.width, max_iter=a.max_iter, want_labels=a.labels)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev[0].record()
    res = km.fit(rows, init, index_base=r0 * a.width, max_iter=a.max_iter, want_labels=a.labels)
    ev[1].record()
    torch.cuda.synchronize()
    fit_ms = ev[0].elapsed_time(ev[1])
    t = torch.tensor([step_ms, ar_ms, fit_ms, hist_ms, compact_ms, hstep_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, ar_ms, fit_ms, hist_ms, compact_ms, hstep_ms = t.tolist()
    if rank == 0:
        peak = 6560.3
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        npix_rank = rows.shape[0] * rows.shape[1]
        ach = 3.0 * npix_rank / (step_ms / 1e3) / 1e9
        line = {"workload": f"per-pixel k-means, one {a.width}x{a.height} image, K={a.k}, rows sharded over {world} GPU(s)",
                "form": "per-pixel (rows re-read every iteration)" if a.per_pixel else
                        "colour histogram (rows read once; iterations over this rank's share of the distinct colours)",
                "labels_in_fit": bool(a.labels), "synthetic": a.synth, "n_gpus": world, "iterations": res.iters, "fit_ms": fit_ms, "ms_per_iteration": fit_ms / max(1, res.iters),
                "iterations_per_sec": res.iters / (fit_ms / 1e3),
                "assign_kernel_ms": step_ms, "allreduce_ms": ar_ms, "allreduce_share": ar_ms / (step_ms + ar_ms),
                "roofline": {"bound": "hbm", "kernel": "k_pixels_step", "algorithmic_bytes_per_launch": 3 * npix_rank,
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak},
                "histogram": None if a.per_pixel else {
                    "count_table_ms": hist_ms, "allreduce_and_compact_ms": compact_ms, "step_over_colours_ms": hstep_ms,
                    "distinct_colours_rank0": n_colours,
                    "count_table_gbs": 3.0 * npix_rank / (hist_ms / 1e3) / 1e9,
                    "count_table_roofline_frac": 3.0 * npix_rank / (hist_ms / 1e3) / 1e9 / peak},
                "centres_sha256": hashlib.sha256(res.centers.cpu().numpy().tobytes()).hexdigest(),
                "counts": res.sums_counts[:, 3].cpu().tolist()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
