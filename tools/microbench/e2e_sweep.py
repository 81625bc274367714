import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
from low_level_feature_extraction_b200.synth import design_image
B,H,W=256,1080,1920
base=np.stack([design_image(H,W,s) for s in range(4)])
host_in=torch.from_numpy(base).repeat(B//4,1,1,1).contiguous().pin_memory()
for hc,ns in ((16,3),(16,4),(32,3),(12,4)):
    an=BatchAnalyzer(0,H,W,BatchConfig(host_chunk=hc,host_streams=ns))
    out=an.alloc_host_outputs(B)
    an.run_host(host_in,out); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(3): an.run_host(host_in,out)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
    print(f"host_chunk={hc} streams={ns}: {dt*1e3:.1f} ms/step  {B/dt:.0f} img/s", flush=True)
    del an, out
    torch.cuda.empty_cache()
