import sys; sys.path.insert(0,'.')
import numpy as np, torch
import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.synth import design_image
eng=pkg.engine(0)
def run(name, batch):
    d=torch.from_numpy(batch).cuda()
    g=eng.gray_blur5(d)
    eng.canny(g); torch.cuda.synchronize()
    eng.ctx.profile_begin()
    for _ in range(3): eng.canny(g)
    p=eng.ctx.profile_end()
    print(name, {k: round(v['ms']/3,3) for k,v in p.items()})
run('design32', np.stack([design_image(1080,1920,s) for s in range(8)]*4))
run('flat32', np.full((32,1080,1920,3),200,np.uint8))
one=design_image(1080,1920,0)
run('design1', one[None])
run('noise8', np.random.default_rng(0).integers(0,256,(8,1080,1920,3),dtype=np.uint8))
