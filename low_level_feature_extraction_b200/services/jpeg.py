"""`cv2.imdecode(buf, cv2.IMREAD_COLOR)` for baseline and progressive JPEG input with everything after the entropy decoding on the GPU
(reference call sites: app/services/analyze/utils.py:108-109, image_processor.py:62-66, :208-211).

The Huffman-coded segment is one serial bit-level decode and runs on the calling host thread inside the library
(csrc/k_jpeg.cu), writing the quantised coefficients into pinned memory; dequantisation, libjpeg-turbo's islow IDCT, the
"fancy" chroma up-sampling and the YCbCr -> BGR conversion run on the device, bit for bit what OpenCV's libjpeg-turbo
produces.  Files outside the subset (arithmetic-coded, 12-bit, CMYK / Adobe-marked, sequential files with several scans, incomplete
progressions, unusual sampling, an Exif segment whose orientation OpenCV would apply) and files the decoder finds damaged return None and the
caller hands the buffer to `cv2.imdecode`, so OpenCV keeps deciding what those decode to."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _runtime

SIGNATURE = b"\xff\xd8"


def decode(buf) -> np.ndarray | None:
    from .._native import LLFE_E_INVALID, LLFE_E_UNSUPPORTED, LlfeError, load_library

    b = bytes(buf) if not isinstance(buf, (bytes, bytearray)) else buf
    info = (C.c_int32 * 2)()
    if load_library().llfe_jpeg_info(b, len(b), info) != 0:
        return None
    w, h = int(info[0]), int(info[1])
    out = np.empty((h, w, 3), np.uint8)
    with _runtime.lock():
        try:
            _runtime.context().call("llfe_jpeg_decode_host", b, len(b), h, w, out)
        except LlfeError as e:
            if e.code in (LLFE_E_INVALID, LLFE_E_UNSUPPORTED):
                return None
            raise
    return out
