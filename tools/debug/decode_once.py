"""Each kernel of the decode / thumbnail row once, for ncu (148 x 1080p PNG streams; one 4K -> 1080p LANCZOS resample)."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from PIL import Image
import low_level_feature_extraction_b200 as pkg
from low_level_feature_extraction_b200.services import png
from low_level_feature_extraction_b200.synth import design_image
eng = pkg.engine(0)
img = design_image(1080, 1920, 0)
b = io.BytesIO(); Image.fromarray(img[:, :, ::-1]).save(b, "PNG")          # Pillow: adaptive filters
info = png.parse(b.getvalue())
stream = np.frombuffer(png.inflate(info), np.uint8)
host = torch.from_numpy(np.tile(stream, (148, 1)))
for _ in range(2):
    out, st = eng.png_reconstruct(host.cuda(), 1080, 1920, info.color_type, info.bit_depth)
big = torch.from_numpy(design_image(2160, 3840, 1)).cuda()
for _ in range(2):
    t = eng.pil_resample_lanczos(big, 1080, 1920)
    r = eng.pil_reduce(big, 2, 2)
torch.cuda.synchronize()
assert np.array_equal(out[147].cpu().numpy(), img)
print("ok")
