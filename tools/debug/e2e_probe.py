"""Time BatchAnalyzer.run_host (256 x 1080p, pinned host in / out) under a few options."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image

B, H, W = 256, 1080, 1920
base = np.stack([design_image(H, W, s) for s in range(8)])
host = torch.from_numpy(np.concatenate([base] * (B // 8))).pin_memory()
for name, cfgkw, opts in [("default (packed masks)", {}, {}), ("u8 masks over PCIe", {"packed_masks": False}, {}),
                          ("packed, 2 threads x 2 workers", {"expand_threads": 2, "expand_workers": 2}, {}),
                          ("packed, 8 threads x 3 workers", {"expand_threads": 8, "expand_workers": 3}, {}),
                          ("packed, chunk 32", {"host_chunk": 32}, {})]:
    an = BatchAnalyzer(0, H, W, BatchConfig(**cfgkw))
    for e in an.engines:
        for k, v in opts.items():
            e.ctx.set_option(k, v)
    out = an.alloc_host_outputs(B)
    an.run_host(host, out)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        an.run_host(host, out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print(name, round(dt * 1e3, 2), "ms per 256 ->", round(B / dt, 1), "img/s", flush=True)
    del an, out
    torch.cuda.empty_cache()
