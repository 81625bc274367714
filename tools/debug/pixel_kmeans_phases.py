#!/usr/bin/env python
"""Where the time of one `PixelKMeans.fit` goes (config 5, one GPU): per-kernel CUDA-event times from the
library's profiler plus host wall-clock per phase.  Debug aid, not a benchmark."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch

from run_pixel_kmeans import synth_rows, synth_rows_photo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--synth", default="design")
    ap.add_argument("--per-pixel", action="store_true")
    a = ap.parse_args()
    import low_level_feature_extraction_b200 as pkg
    from low_level_feature_extraction_b200.dist import PixelKMeans

    eng = pkg.engine(0)
    dev = torch.device("cuda", 0)
    f = synth_rows_photo if a.synth == "photo" else synth_rows
    rows = torch.cat([f(s, min(a.size, s + 1024), a.size, dev) for s in range(0, a.size, 1024)], dim=0)
    g = torch.Generator().manual_seed(42)
    pos = torch.randint(0, a.size * a.size, (a.k,), generator=g)
    init = torch.stack([rows[p // a.size, p % a.size].flip(0).to(torch.float32) for p in pos.tolist()])
    km = PixelKMeans(eng, histogram=not a.per_pixel)
    km.fit(rows, init)
    torch.cuda.synchronize()
    for labels in (False, True):
        t0 = time.perf_counter()
        res = km.fit(rows, init, want_labels=labels)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.ctx.profile_begin()
        km.fit(rows, init, want_labels=labels)
        prof = eng.ctx.profile_end()
        print(json.dumps({"labels": labels, "iters": res.iters, "wall_ms": (t1 - t0) * 1e3, "profile": prof}))


if __name__ == "__main__":
    main()
