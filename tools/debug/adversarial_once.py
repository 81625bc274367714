import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from low_level_feature_extraction_b200.services import ColorExtractor
img = np.random.default_rng(2024).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
px = img.reshape(-1, 3)
ColorExtractor.set_rng_seed(77)
ColorExtractor._get_dominant_colors(px, 5)
ColorExtractor.set_rng_seed(77)
t = time.perf_counter(); c, l = ColorExtractor._get_dominant_colors(px, 5); print("adversarial frame: %.2f s" % (time.perf_counter() - t))
