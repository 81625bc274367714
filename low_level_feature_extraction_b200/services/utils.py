"""Drop-in for the hot-path part of app/services/analyze/utils.py:
`validate_and_preprocess_image` (:90-152).  PNG input: the host inflates the IDAT stream, scanline
reconstruction and the conversion to BGR run on the GPU (services/png.py, csrc/k_png.cu); every other
format goes through cv2.imdecode like the reference (sequential entropy decoding).  The `auto`
INTER_AREA down-scale -- the mode the endpoint hard-codes (endpoints/analyze.py:90) --
and the `performance` (INTER_LINEAR) and `high_quality` (INTER_LANCZOS4) down-scales run on the GPU.  Download and response assembly (:31-87, :155-214) are network /
HTTP glue outside the path."""
from __future__ import annotations

import logging
from enum import Enum

import numpy as np

from .image_processor import resize_area, resize_lanczos4, resize_linear
from .png import imdecode_color

logger = logging.getLogger(__name__)

try:
    from fastapi import HTTPException, status

    _BAD_REQUEST = status.HTTP_400_BAD_REQUEST
except Exception:  # pragma: no cover - fastapi is optional outside the web app
    class HTTPException(Exception):
        def __init__(self, status_code: int, detail: str = ""):
            super().__init__(detail)
            self.status_code = status_code
            self.detail = detail

    _BAD_REQUEST = 400


class PreprocessingMode(str, Enum):
    NONE = "none"
    AUTO = "auto"
    HIGH_QUALITY = "high_quality"
    PERFORMANCE = "performance"


async def validate_and_preprocess_image(image_bytes: bytes, request_id: str, preprocessing: str) -> np.ndarray:
    """bytes -> (H, W, 3) uint8 BGR; any failure -> HTTPException(400)."""
    try:
        image = imdecode_color(image_bytes)                                            # utils.py:108-109
        if image is None:
            raise HTTPException(status_code=_BAD_REQUEST,
                                detail="Failed to decode image. The file may be corrupted or in an unsupported format.")
        if preprocessing == "none":
            pass
        elif preprocessing == "auto":
            max_dim = 2000
            h, w = image.shape[:2]
            if max(h, w) > max_dim:
                scale = max_dim / max(h, w)
                image = resize_area(image, int(w * scale), int(h * scale))            # utils.py:125-127
        elif preprocessing == "high_quality":
            max_dim = 4000
            h, w = image.shape[:2]
            if max(h, w) > max_dim:
                scale = max_dim / max(h, w)
                image = resize_lanczos4(image, int(w * scale), int(h * scale))        # utils.py:133-135
        elif preprocessing == "performance":
            max_dim = 1000
            h, w = image.shape[:2]
            if max(h, w) > max_dim:
                scale = max_dim / max(h, w)
                image = resize_linear(image, int(w * scale), int(h * scale))          # utils.py:141-143
        return image
    except Exception as e:
        logger.error(f"Error in validate_and_preprocess_image: {str(e)}", exc_info=True)
        raise HTTPException(status_code=_BAD_REQUEST, detail=f"Image validation or preprocessing failed: {str(e)}")
