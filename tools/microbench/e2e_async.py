"""e2e throughput of BatchAnalyzer.run_host (one synchronous call per batch) against run_host_async (two batches in
flight), alternating blocks of 6 batches on the same box; masks + palettes (no host palette tail)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image
B = 256
base = np.stack([design_image(1080, 1920, s) for s in range(16)])
host_in = torch.from_numpy(np.concatenate([base] * (B // 16))).pin_memory()
an = BatchAnalyzer(0, 1080, 1920, BatchConfig())
outs = [an.alloc_host_outputs(B) for _ in range(2)]
an.run_host(host_in, outs[0]); an.run_host_async(host_in, outs[1]).result(); an.run_host_async(host_in, outs[0]).result()
torch.cuda.synchronize()
def sync_block(steps=6):
    t = time.perf_counter()
    for i in range(steps): an.run_host(host_in, outs[i % 2])
    return B * steps / (time.perf_counter() - t)
def async_block(steps=6):
    t = time.perf_counter(); prev = None
    for i in range(steps):
        c = an.run_host_async(host_in, outs[i % 2])
        if prev is not None: prev.result()
        prev = c
    prev.result()
    return B * steps / (time.perf_counter() - t)
rs, ra = [], []
for _ in range(8):
    rs.append(sync_block()); ra.append(async_block())
print("sync  blocks", [round(v) for v in rs], "median", round(float(np.median(rs))))
print("async blocks", [round(v) for v in ra], "median", round(float(np.median(ra))))
