import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.synth import design_image
eng=pkg.engine(0)
n=int(sys.argv[1]) if len(sys.argv)>1 else 32
base=np.stack([design_image(1080,1920,s) for s in range(4)])
d=torch.from_numpy(base).cuda().repeat((n+3)//4,1,1,1)[:n].contiguous()
for _ in range(3):
    out=eng.pipeline(d, seed=1, max_unique=1<<16)
    eng.kmeans_unique(out["keys"], out["count"], 5, [1000+i for i in range(n)])
torch.cuda.synchronize()
print("ok", int(out["count"][0]))
