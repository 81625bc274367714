"""What tools/sanitize.sh runs under compute-sanitizer: the smoke path (fused pipeline with injected noise +
k-means on two 135x240 images) and one device-noise pipeline call with k-means, hysteresis cluster kernel and
bitmap compaction on 4 images of 270x480 (small on purpose: the tools slow kernels down 10-100x)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image

g.smoke()
imgs = torch.from_numpy(np.stack([design_image(270, 480, s) for s in range(4)])).cuda()
an = BatchAnalyzer(0, 270, 480, BatchConfig())
out = an.run_device(imgs)
torch.cuda.synchronize()
print("pipeline ok", out["k_used"].cpu().tolist(), out["count"].cpu().tolist())
