#!/usr/bin/env python
"""Latency of the decode row on one 1080p design PNG: cv2.imdecode (the reference's call) against
services.png (host inflate + device reconstruction), and ImageProcessor.auto_process_image against the reference's
call sequence (cv2.imdecode + Pillow LANCZOS thumbnail) on a 4K PNG."""
import io
import json
import os
import sys
import time
import zlib

import cv2
import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from low_level_feature_extraction_b200.services import png  # noqa: E402
from low_level_feature_extraction_b200.services.image_processor import ImageProcessor, pil_thumbnail_lanczos  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image  # noqa: E402


def med(f, reps=7):
    f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        t.append(time.perf_counter() - t0)
    return float(np.median(t)) * 1e3


def main():
    out = {"unit": "ms (median)"}
    img = design_image(1080, 1920, 0)
    files = {"opencv": cv2.imencode(".png", img)[1].tobytes()}
    b = io.BytesIO()
    Image.fromarray(img[:, :, ::-1]).save(b, "PNG")
    files["pillow"] = b.getvalue()
    for name, buf in files.items():
        info = png.parse(buf)
        stream = png.inflate(info)
        dst = np.empty((1080, 1920, 3), np.uint8)
        from low_level_feature_extraction_b200.services import _runtime
        ctx = _runtime.context()
        assert np.array_equal(png.decode(buf), cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR))
        out[f"1080p_{name}"] = {
            "file_bytes": len(buf),
            "cv2_imdecode": med(lambda: cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)),
            "services_png_decode": med(lambda: png.decode(buf)),
            "of_which_host_parse_crc": med(lambda: png.parse(buf)),
            "of_which_host_inflate": med(lambda: png.inflate(info)),
            "of_which_device_call_with_copies": med(lambda: ctx.call("llfe_png_reconstruct_host", stream, 1080, 1920,
                                                                     info.color_type, info.bit_depth, None, 0, dst)),
        }
    from low_level_feature_extraction_b200.services import jpeg
    for q, prog in ((95, 0), (75, 0), (90, 1)):
        jb = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_PROGRESSIVE, prog])[1].tobytes()
        assert np.array_equal(jpeg.decode(jb), cv2.imdecode(np.frombuffer(jb, np.uint8), cv2.IMREAD_COLOR))
        out[f"1080p_jpeg_q{q}" + ("_progressive" if prog else "")] = {
            "file_bytes": len(jb),
            "cv2_imdecode": med(lambda: cv2.imdecode(np.frombuffer(jb, np.uint8), cv2.IMREAD_COLOR)),
            "services_jpeg_decode": med(lambda: jpeg.decode(jb)),
        }
    big = design_image(2160, 3840, 1)
    buf = cv2.imencode(".png", big)[1].tobytes()

    def ref():
        image = cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR)
        pil_image = Image.fromarray(cv2.cvtColor(image, cv2.COLOR_BGR2RGB))
        pil_image.thumbnail((1920, 1080), Image.Resampling.LANCZOS)
        return cv2.cvtColor(np.array(pil_image), cv2.COLOR_RGB2BGR)

    assert np.array_equal(ref(), ImageProcessor.auto_process_image(buf))

    def pil_only():
        im = Image.fromarray(big)
        im.thumbnail((1920, 1080), Image.Resampling.LANCZOS)
        return np.array(im)

    out["4k_auto_process_image"] = {
        "reference_call_sequence": med(ref, 5),
        "drop_in": med(lambda: ImageProcessor.auto_process_image(buf), 5),
        "pillow_thumbnail_only": med(pil_only, 5),
        "device_thumbnail_only_with_copies": med(lambda: pil_thumbnail_lanczos(big, 1920, 1080), 5),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
