"""Drop-in for the hot-path part of the reference's FontDetector:
`preprocess_image` (app/services/analyze/font_detector.py:16-37).  Contour
heuristics (:39-170) are host code outside the path."""
from __future__ import annotations

import numpy as np

from . import _runtime


class FontDetector:
    @staticmethod
    def preprocess_image(image: np.ndarray) -> np.ndarray:
        """BGR -> gray -> adaptiveThreshold(GAUSSIAN_C, BINARY_INV, 11, 2)."""
        img = _runtime.as_bgr_u8(image)
        h, w = img.shape[:2]
        mask = np.empty((h, w), np.uint8)
        with _runtime.lock():
            _runtime.context().call("llfe_font_mask_host", img, h, w, mask)
        return mask
