import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, cv2
from PIL import Image
import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.services import png
from low_level_feature_extraction_b200.synth import design_image
eng = pkg.engine(0)
img = design_image(1080, 1920, 0)
for name, buf in (("opencv", cv2.imencode(".png", img)[1].tobytes()),):
    info = png.parse(buf)
    stream = np.frombuffer(png.inflate(info), np.uint8)
    for n in (1, 32, 148):
        host = torch.from_numpy(np.tile(stream, (n, 1)))
        for rep in range(3):
            d = host.cuda()
            torch.cuda.synchronize()
            eng.ctx.profile_begin()
            out, st = eng.png_reconstruct(d, 1080, 1920, info.color_type, info.bit_depth)
            torch.cuda.synchronize()
            prof = eng.ctx.profile_end()
        assert np.array_equal(out[n - 1].cpu().numpy(), img)
        print(name, n, prof)
