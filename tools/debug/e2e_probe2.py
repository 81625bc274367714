import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
class A: pass
dev = torch.device("cuda", 0)
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
B,H,W=256,1080,1920
for distinct in (8, 64):
    batch = bench.device_batch(dev, 0, B, H, W, distinct)
    an = BatchAnalyzer(0, H, W, BatchConfig())
    host_in = torch.empty(tuple(batch.shape), dtype=torch.uint8).pin_memory(); host_in.copy_(batch)
    out = an.alloc_host_outputs(B)
    an.run_host(host_in, out); torch.cuda.synchronize()
    for rep in range(3):
        t=time.perf_counter(); res=an.run_host(host_in, out); torch.cuda.synchronize(); dt=time.perf_counter()-t
        print("distinct",distinct,"run_host ms", round(dt*1e3,2), "overflow", int((res["count"]>65536).sum()), flush=True)
    t=time.perf_counter(); p=an.palettes(res); print("tail ms", round((time.perf_counter()-t)*1e3,2))
    del an, batch, host_in, out
