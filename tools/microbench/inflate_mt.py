#!/usr/bin/env python
"""llfe_inflate_zlib_mt (csrc/h_inflate.cu: several decoders on one deflate stream) on the IDAT streams of 1080p / 4K
design PNGs written by OpenCV and Pillow, per number of decoders, next to zlib.  Host only (no GPU needed).

    python tools/microbench/inflate_mt.py > profiles/inflate_mt_<where>.json
"""
import ctypes as C
import io
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import cv2
import numpy as np
from PIL import Image

from low_level_feature_extraction_b200._native import load_library
from low_level_feature_extraction_b200.services import png
from low_level_feature_extraction_b200.synth import design_image


def best(f, reps):
    b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        b = min(b, time.perf_counter() - t0)
    return round(b * 1e3, 2)


def main():
    lib = load_library()
    out = {"unit": "ms (best of 9, after a warm-up pass over every thread count)", "host_cpus": os.cpu_count()}
    for (h, w) in ((1080, 1920), (2160, 3840)):
        img = design_image(h, w, 0)
        files = {"opencv": cv2.imencode(".png", img)[1].tobytes()}
        b = io.BytesIO()
        Image.fromarray(img[:, :, ::-1]).save(b, "PNG")
        files["pillow"] = b.getvalue()
        for name, buf in files.items():
            info = png.parse(buf)
            n = info.stream_bytes
            ref = zlib.decompress(info.idat)
            dst = np.zeros(n, np.uint8)
            got = C.c_size_t(0)
            rec = {"idat_bytes": len(info.idat), "stream_bytes": n, "zlib": best(lambda: zlib.decompress(info.idat), 5)}
            for rnd in range(2):      # the first pass warms the process up (threads, buffers)
                for th in (1, 2, 3, 4, 6, 8):
                    def run():
                        rc = lib.llfe_inflate_zlib_mt(info.idat, len(info.idat), dst.ctypes.data, n, C.byref(got), th)
                        assert rc == 0 and got.value == n
                    run()
                    assert dst.tobytes() == ref
                    rec[f"decoders_{th}"] = best(run, 9)
            out[f"{w}x{h} {name}"] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
