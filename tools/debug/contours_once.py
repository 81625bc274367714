#!/usr/bin/env python
"""Two single-image calls of the device contour finder on a 1080p design shape mask (the ncu target for the contour
kernels: `ncu --set full -k regex:k_ct_ --launch-skip 7 -c 7 python tools/debug/contours_once.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch

import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.synth import design_image
from oracle import refpath  # the mask only (input of the call under the profiler)

eng = pkg.engine(0)
d = torch.from_numpy(refpath.shape_mask(design_image(1080, 1920, 0))).cuda()
for _ in range(2):
    hdr, pts, cnt = eng.contours_external(d, 200, 1 << 15, 1 << 19)
    torch.cuda.synchronize()
c = cnt.reshape(-1).cpu()
print("contours", int(c[0]), "points", int(c[1]))
