import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.synth import design_image
eng=pkg.engine(0)
n=int(sys.argv[1]) if len(sys.argv)>1 else 8
base=np.stack([design_image(1080,1920,s) for s in range(n)])
d=torch.from_numpy(base).cuda()
out=eng.pipeline(d, seed=1, max_unique=1<<16, shapes=False, shadows=False)
dbg=torch.zeros((n,10,8),dtype=torch.int64,device='cuda')
for _ in range(2): eng.kmeans_unique(out["keys"], out["count"], 5, [1000+i for i in range(n)])
torch.cuda.synchronize()
eng.ctx.set_debug_buffer("kmeans", dbg)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); eng.kmeans_unique(out["keys"], out["count"], 5, [1000+i for i in range(n)]); e1.record(); torch.cuda.synchronize()
print("kmeans ms", e0.elapsed_time(e1), "for", n, "images")
x=dbg.cpu().numpy()
np.set_printoptions(linewidth=200)
print("per (img, attempt): [load, seed, first, rest, final] kcycles; iters, U, total")
for i in range(min(n,2)):
    print(np.concatenate([x[i,:,:5]//1000, x[i,:,5:7], x[i,:,7:]//1000],axis=1))
c=x[:,:,0]; q=(c>>40); f=(c>>20)&0xfffff; ch=c&0xfffff
print("queued/full/changed per attempt (img0):", list(zip(q[0].tolist(), f[0].tolist(), ch[0].tolist())))
print("mean kcycles:", (x[:,:,:5].mean((0,1))/1000).round(1), "mean iters", x[:,:,5].mean(), "mean total kcycles", x[:,:,7].mean()/1000)
