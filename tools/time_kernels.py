"""Per-kernel CUDA-event times of the pipeline on B images of HxW (device resident), for A/B runs of options.

    python tools/time_kernels.py [--batch 256] [--opt shadow_inline=1] [--steps 5] [--workload pipeline]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--distinct", type=int, default=8)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--workload", default="pipeline")
a = ap.parse_args()
B, H, W = a.batch, a.height, a.width
base = torch.from_numpy(np.stack([design_image(H, W, s) for s in range(a.distinct)])).cuda()
batch = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
for i in range(B):
    batch[i] = torch.roll(base[i % a.distinct], shifts=7 * (i // a.distinct), dims=0)
cfg = BatchConfig(colors=a.workload in ("pipeline", "colors"), shapes=a.workload in ("pipeline", "shapes"),
                  shadows=a.workload in ("pipeline", "shadows"))
an = BatchAnalyzer(0, H, W, cfg)
eng = an.engines[0]
for o in a.opt:
    k, v = o.split("=")
    eng.ctx.set_option(k, int(v))
out = an.alloc_outputs(B)
for _ in range(2):
    an.run_device(batch, out)
torch.cuda.synchronize()
eng.ctx.profile_begin()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    an.run_device(batch, out)
e1.record()
torch.cuda.synchronize()
k = eng.ctx.profile_end()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"opts": a.opt, "workload": a.workload, "ms_per_step": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1),
                  "kernels_ms_per_step": {n: round(v["ms"] / a.steps, 4) for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"])}}))
