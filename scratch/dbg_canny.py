import sys; sys.path.insert(0,'.')
import numpy as np, torch
import low_level_feature_extraction_b200 as pkg
from oracle import cvops
eng=pkg.engine(0)
for shape in [(53,37),(64,96),(135,257),(33,2049),(40,64),(40,37),(64,37),(53,64)]:
    r=np.random.default_rng(shape[1]); noise=r.integers(0,256,shape,dtype=np.uint8)
    for nm,src in (('blur',cvops.gaussian_blur5(noise)),('noise',noise)):
        e=cvops.canny(src,50,150)
        o=eng.canny(torch.from_numpy(src).cuda(),50,150).cpu().numpy()
        bad=np.argwhere(o!=e)
        w,s=cvops.canny_nms(src,50,150)
        print(shape,nm,'mismatch',len(bad), 'missing',int(((o==0)&(e==255)).sum()),'extra',int(((o==255)&(e==0)).sum()), 'extra not weak', int(((o==255)&(~w)).sum()), bad[:6].tolist())
