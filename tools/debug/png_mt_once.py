#!/usr/bin/env python
"""services.png.decode with 1 and 4 inflate decoders (llfe_set_option("inflate_threads")) on 1080p and 4K design
PNGs, checked against cv2.imdecode and timed next to it.

    python tools/debug/png_mt_once.py
"""
import io
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import cv2
import numpy as np
from PIL import Image

from low_level_feature_extraction_b200.services import _runtime, png
from low_level_feature_extraction_b200.synth import design_image
from oracle import pngops  # writes the Adam7 test file only


def med(f, reps=5):
    f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        t.append((time.perf_counter() - t0) * 1e3)
    return round(statistics.median(t), 2)


def main():
    ctx = _runtime.context()
    out = {"unit": "ms (median of 5)", "host_cpus": os.cpu_count()}
    files = {}
    img = design_image(1080, 1920, 0)
    files["1080p opencv"] = cv2.imencode(".png", img)[1].tobytes()
    b = io.BytesIO()
    Image.fromarray(img[:, :, ::-1]).save(b, "PNG")
    files["1080p pillow"] = b.getvalue()
    files["4k opencv"] = cv2.imencode(".png", design_image(2160, 3840, 1))[1].tobytes()
    files["1080p adam7"] = pngops.write_png_interlaced(img[:, :, ::-1].astype(np.int64), 2, 8, np.random.default_rng(0))
    for th in (4, 1, 4):       # (the first round also warms the process up)
        ctx.set_option("inflate_threads", th)
        for name, buf in files.items():
            ref = cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)
            assert np.array_equal(png.decode(buf), ref), (name, th)
            out.setdefault(name, {"file_bytes": len(buf)})[f"services_png_decode_{th}_decoders"] = med(lambda: png.decode(buf))
    for name, buf in files.items():
        out[name]["cv2_imdecode"] = med(lambda: cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR), 5)
    ctx.set_option("inflate_threads", 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
