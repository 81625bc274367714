#!/usr/bin/env python
"""Host <-> device copy bandwidth with all ranks copying at once (what bounds `e2e` at N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/microbench/pcie_multi.py

For each rank: the GPU's PCI bus id, NUMA node and local CPU list (sysfs), the process' CPU affinity;
then aggregate H2D, D2H and bidirectional GB/s with every rank copying concurrently -- first with the
affinity the launcher gave us, then after binding the process to the GPU's local CPUs and re-allocating the
pinned buffers (first touch on the local node).  Rank 0 prints one JSON line.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from low_level_feature_extraction_b200.numa import gpu_locality, bind_to_gpu  # noqa: E402


def measure(dev, nbytes, world):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out = torch.empty(nbytes * 2 // 3, dtype=torch.uint8).pin_memory()
    h_out.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(nbytes * 2 // 3, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(f, reps=4):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            f()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = min(best, float(dt.item()))
        return best

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    t_h2d = run(lambda: d_in.copy_(h_in, non_blocking=True))
    t_d2h = run(lambda: h_out.copy_(d_out, non_blocking=True))
    t_both = run(both)
    return {"h2d_gbs_total": world * nbytes / t_h2d / 1e9, "d2h_gbs_total": world * (nbytes * 2 // 3) / t_d2h / 1e9,
            "both_h2d_gbs_total": world * nbytes / t_both / 1e9}


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    loc = gpu_locality(local)
    info = {"rank": rank, **loc, "affinity_before": len(os.sched_getaffinity(0))}
    before = measure(dev, nbytes, world)
    bound = bind_to_gpu(local)
    info["bound"] = bound
    info["affinity_after"] = len(os.sched_getaffinity(0))
    after = measure(dev, nbytes, world)
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "cpus": os.cpu_count(), "default_affinity": before, "bound_to_gpu_node": after,
                          "ranks": infos}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
