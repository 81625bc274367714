#!/usr/bin/env python
"""Which part of `BatchAnalyzer.run_host` bounds the end-to-end rate: the same staged loop with
(a) only the host->device copies, (b) copies in both directions, (c) copies in + kernels, (d) everything."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch

import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
import numpy as np

base = np.stack([design_image(1080, 1920, s) for s in range(8)])
imgs = torch.from_numpy(np.concatenate([base] * (n // 8))).pin_memory()
an = BatchAnalyzer(0, 1080, 1920, BatchConfig())
host_out = an.alloc_host_outputs(n)
real_run = an.run_device
out_keys = ("shape_mask", "shadow_mask", "shadow_sums", "centers", "count", "k_used", "cluster_sizes")


def timed(label, run_device, d2h):
    an.run_device = run_device
    ho = host_out if d2h else {k: v for k, v in host_out.items() if k not in out_keys}
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        an.run_host(imgs, ho)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return {"variant": label, "ms": best * 1e3, "images_per_sec": n / best}


res = [timed("full", real_run, True)]
res.append(timed("h2d_only", lambda *a, **k: None, False))
res.append(timed("h2d_d2h", lambda *a, **k: None, True))
res.append(timed("h2d_kernels", real_run, False))
res.append(timed("full_again", real_run, True))
print(json.dumps({"n": n, "h2d_bytes": imgs.numel(), "results": res}))
