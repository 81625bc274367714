"""Drop-in for the hot-path part of the reference's TextExtractor:
`preprocess_image` (app/services/analyze/text_extractor.py:15-46).  The OCR half
of that module (pytesseract, :48-346) is out of scope (external process)."""
from __future__ import annotations

import cv2
import numpy as np

from . import _runtime


class TextExtractor:
    @classmethod
    def preprocess_image(cls, image: np.ndarray) -> np.ndarray:
        """gray -> (INTER_CUBIC upscale of tiny images) -> Otsu -> invert if mostly white."""
        color = len(image.shape) > 2 and image.shape[2] > 1
        height, width = image.shape[:2]
        if height < 30 or width < 100 or not color:
            # tiny images take the reference's INTER_CUBIC upscale first (text_extractor.py:33-37), which is
            # outside the hot path; gray inputs skip the colour conversion.  Do those
            # two host steps as the reference does, then run Otsu + invert on the GPU.
            gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if color else image
            if height < 30 or width < 100:
                scale = max(2, 300 / width, 100 / height)
                gray = cv2.resize(gray, None, fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
            gray = np.ascontiguousarray(gray)
            img3 = np.repeat(gray[..., None], 3, axis=2)   # BGR2GRAY of (g,g,g) is g: (3735+19235+9798) g + 16384 >> 15
        else:
            img3 = _runtime.as_bgr_u8(image)
        h, w = img3.shape[:2]
        mask = np.empty((h, w), np.uint8)
        thr = np.zeros(1, np.int32)
        with _runtime.lock():
            _runtime.context().call("llfe_text_mask_host", img3, h, w, mask, thr)
        return mask
