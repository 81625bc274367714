"""Drop-in for the reference's ShadowAnalyzer (the /extract-shadows service).

Mirrors app/services/__pycache__/shadow_analyzer.cpython-312.pyc (src L4-31).
gray -> blur5 -> adaptive threshold -> masked mean runs fused on the GPU; only
the final three-way comparison is host Python, as in the reference (L24-31).
"""
from __future__ import annotations

import numpy as np

from . import _runtime


class ShadowAnalyzer:
    @staticmethod
    def preprocess_image(image: np.ndarray) -> np.ndarray:
        """BGR -> gray -> GaussianBlur((5,5),0)   (src L5-10)."""
        img = _runtime.as_bgr_u8(image)
        h, w = img.shape[:2]
        blurred = np.empty((h, w), np.uint8)
        sc = np.zeros(2, np.uint64)
        with _runtime.lock():
            _runtime.context().call("llfe_shadow_mask_host", img, h, w, None, blurred, sc)
        return blurred

    @staticmethod
    def shadow_mask(image: np.ndarray):
        """-> (threshold mask (H,W) u8, sum of blurred pixels under the mask, their count)."""
        img = _runtime.as_bgr_u8(image)
        h, w = img.shape[:2]
        mask = np.empty((h, w), np.uint8)
        sc = np.zeros(2, np.uint64)
        with _runtime.lock():
            _runtime.context().call("llfe_shadow_mask_host", img, h, w, mask, None, sc)
        return mask, int(sc[0]), int(sc[1])

    @staticmethod
    def analyze_shadow_level(image: np.ndarray) -> str:
        """'Low' | 'Moderate' | 'High'   (src L12-31)."""
        _, total, count = ShadowAnalyzer.shadow_mask(image)
        return ShadowAnalyzer.level_from_sums(total, count)

    @staticmethod
    def level_from_sums(total: int, count: int) -> str:
        """src L19-31 from the exact sum / count of the blurred pixels under the threshold mask."""
        if count == 0:
            return "Low"
        avg_darkness = 255 - total / count   # np.mean of the masked pixels: exact integer sum / count in float64
        if avg_darkness < 30:
            return "Low"
        elif avg_darkness < 60:
            return "Moderate"
        return "High"
