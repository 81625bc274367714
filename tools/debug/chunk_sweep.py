import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image
n = 256
base = np.stack([design_image(1080, 1920, s) for s in range(32)])
bgr = torch.from_numpy(np.concatenate([base] * 8)).cuda()
ba = BatchAnalyzer(0, 1080, 1920, BatchConfig())
out = ba.alloc_outputs(n)
for chunk in (64, 96, 128, 256, 64, 32):
    ba.engines[0].ctx.set_option("chunk", chunk)
    for _ in range(3): ba.run_device(bgr, out, resolve=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ba.run_device(bgr, out, resolve=False)
    e1.record(); torch.cuda.synchronize()
    print("chunk", chunk, "ms/step", round(e0.elapsed_time(e1) / 10, 3))
