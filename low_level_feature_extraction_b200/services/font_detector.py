"""Drop-in for the hot-path part of the reference's FontDetector:
`preprocess_image` (app/services/analyze/font_detector.py:16-37) and
`detect_text_regions` (:40-68, `cv2.findContours` + `cv2.boundingRect` per
contour: device contour finder, only the bounding boxes come back).  The font
heuristics (:70-170) are host code outside the path."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import _runtime
from .. import contours as _contours


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

    @staticmethod
    def detect_text_regions(image: np.ndarray) -> List[Tuple[int, int, int, int]]:
        """Preprocessed binary image -> [(x, y, w, h)] of the external contours that look like text (:51-68)."""
        import cv2

        if not isinstance(image, np.ndarray) or image.ndim != 2 or image.dtype != np.uint8 or image.size == 0:
            raise cv2.error("findContours: (-210:Unsupported format or combination of formats) [Start]FindContours "
                            "supports only CV_8UC1 images when mode != CV_RETR_FLOODFILL")
        with _runtime.lock():
            headers, _ = _contours.find_external_host(_runtime.context(), image, _contours.NO_POINTS, want_points=False)
        text_regions = []
        for x, y, w, h in _contours.bounding_rects(headers):
            aspect_ratio = w / float(h)
            if 0.1 < aspect_ratio < 15 and h > 8:
                text_regions.append((x, y, w, h))
        return text_regions
