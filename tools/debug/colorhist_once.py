#!/usr/bin/env python
"""Launch each colour-histogram kernel of config 5 once or twice on a photo-like image (for an ncu capture)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch

from run_pixel_kmeans import synth_rows_photo  # noqa: E402

import low_level_feature_extraction_b200 as pkg  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = pkg.engine(0)
dev = torch.device("cuda", 0)
rows = torch.cat([synth_rows_photo(s, min(size, s + 1024), size, dev) for s in range(0, size, 1024)], dim=0)
init = torch.rand((16, 3), device=dev) * 255
hist = torch.zeros((1 << 24,), dtype=torch.int32, device=dev)
for _ in range(2):
    hist.zero_()
    eng.pixels_histogram(rows, hist)
keys, counts = eng.histogram_compact(hist)
sums = torch.zeros((16, 4), dtype=torch.int64, device=dev)
lab = torch.empty((keys.numel(),), dtype=torch.uint8, device=dev)
for _ in range(2):
    eng.kmeans_hist_step(keys, counts, init, sums, lab)
lut = torch.zeros((1 << 24,), dtype=torch.uint8, device=dev)
eng.hist_labels_to_lut(keys, lab, lut)
out = torch.empty((size * size,), dtype=torch.uint8, device=dev)
for _ in range(2):
    eng.pixels_lookup(rows, lut, out)
torch.cuda.synchronize()
print("colours", keys.numel(), "pixels", size * size)
