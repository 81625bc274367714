#!/usr/bin/env python
"""Config 5 on N GPUs of one node: `PixelKMeans.fit` with the fused NVLink exchange (llfe_kmeans_update_p2p) next to the
ncclAllReduce + update loop -- identical centres / iteration counts, time of the whole fit for both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_check.py [--size 16384]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch
import torch.distributed as dist

from run_pixel_kmeans import synth_rows, synth_rows_photo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--synth", default="design")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import low_level_feature_extraction_b200 as pkg
    from low_level_feature_extraction_b200.dist import PixelKMeans, shard_range

    eng = pkg.engine(local)
    r0, r1 = shard_range(a.size, rank, world)
    gen = synth_rows_photo if a.synth == "photo" else synth_rows
    rows = torch.cat([gen(s, min(r1, s + 1024), a.size, dev) for s in range(r0, r1, 1024)], dim=0)
    g = torch.Generator().manual_seed(42)
    pos = torch.randint(0, a.size * a.size, (a.k,), generator=g)
    init = torch.zeros((a.k, 3), dtype=torch.float32, device=dev)
    for j, p in enumerate(pos.tolist()):
        y, x = divmod(p, a.size)
        if r0 <= y < r1:
            init[j] = rows[y - r0, x].flip(0).to(torch.float32)
    dist.all_reduce(init)
    out = {}
    for name, p2p in (("nccl_allreduce", False), ("p2p_fused", True)):
        km = PixelKMeans(eng, p2p=p2p)
        res = km.fit(rows, init, index_base=r0 * a.size)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            res = km.fit(rows, init, index_base=r0 * a.size)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"fit_ms": float(t.item()), "iterations": res.iters,
                     "centres_sha256": hashlib.sha256(res.centers.cpu().numpy().tobytes()).hexdigest()}
    if rank == 0:
        print(json.dumps({"world": world, "size": a.size, "k": a.k, "synth": a.synth,
                          "identical": out["nccl_allreduce"]["centres_sha256"] == out["p2p_fused"]["centres_sha256"]
                                       and out["nccl_allreduce"]["iterations"] == out["p2p_fused"]["iterations"], **out}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
