import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.synth import design_image
eng=pkg.engine(0)
n=int(sys.argv[1]) if len(sys.argv)>1 else 8
base=np.stack([design_image(1080,1920,s) for s in range(n)])
d=torch.from_numpy(base).cuda()
dbg=torch.zeros((n,16,8),dtype=torch.int64,device='cuda')
for _ in range(2): eng.shape_mask(d)
torch.cuda.synchronize()
eng.ctx.set_debug_buffer("hysteresis", dbg)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); eng.shape_mask(d); e1.record(); torch.cuda.synchronize()
print("shape_mask ms", e0.elapsed_time(e1))
x=dbg.cpu().numpy()[:,:8]
np.set_printoptions(linewidth=200)
print("per CTA [load, local, sync, conv_total, expand] kcycles; iters, rounds, has_cand")
for i in range(min(n,4)):
    print("img",i); print(np.concatenate([x[i,:,:5]//1000, x[i,:,5:]],axis=1))
print("mean kcycles:", (x[:,:,:5].mean((0,1))/1000).round(1), "mean iters", x[:,:,5].mean(), "max", x[:,:,5].max(), "rounds mean", x[:,:,6].mean(), "max", x[:,:,6].max())
