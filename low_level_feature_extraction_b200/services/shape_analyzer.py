"""Drop-in for the reference's ShapeAnalyzer (the /extract-shapes service).

Mirrors app/services/__pycache__/shape_analyzer.cpython-312.pyc (source lines
L5-189): same class, method names, arguments, return shapes and error
behaviour.  `preprocess_image` (gray -> blur -> Canny -> dilate, src L6-30) and the
contour finder (`cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)`, src
L76 / L140; csrc/k_contours.cu) run on the GPU through libllfe.so: for
`analyze_shapes` / `extract_shapes` the mask never leaves the device, only the
vertices of the contours that pass the reference's area threshold come back.
The per-contour polygon geometry (approxPolyDP, convexHull, arcLength; src
L142-181) stays the reference's own cv2 calls on those few points.
"""
from __future__ import annotations

from typing import Any, Dict, List

import cv2
import numpy as np

from . import _runtime
from .. import contours as _contours


class ShapeAnalyzer:
    CANNY_LOW = 50
    CANNY_HIGH = 150

    @staticmethod
    def preprocess_image(image: np.ndarray) -> np.ndarray:
        """(H,W,3) u8 BGR -> (H,W) u8 dilated edge mask in {0,255}  (src L6-30)."""
        img = _runtime.as_bgr_u8(image)
        h, w = img.shape[:2]
        mask = np.empty((h, w), np.uint8)
        with _runtime.lock():
            _runtime.context().call("llfe_shape_mask_host", img, h, w, ShapeAnalyzer.CANNY_LOW,
                                    ShapeAnalyzer.CANNY_HIGH, mask)
        return mask

    @staticmethod
    def detect_border_radius(contour, epsilon_factor: float = 0.02) -> float:
        """src L32-61."""
        epsilon = epsilon_factor * cv2.arcLength(contour, True)
        approx = cv2.approxPolyDP(contour, epsilon, True)
        if len(approx) > 4:
            hull = cv2.convexHull(contour)
            hull_area = cv2.contourArea(hull)
            contour_area = cv2.contourArea(contour)
            if hull_area > 0:
                area_ratio = 1 - contour_area / hull_area
                border_radius = area_ratio * 50.0
                return max(0.0, border_radius)
        return 0.0

    @staticmethod
    def _classify(contour) -> str:
        epsilon = 0.04 * cv2.arcLength(contour, True)
        approx = cv2.approxPolyDP(contour, epsilon, True)
        shape_type = "unknown"
        if len(approx) == 3:
            shape_type = "triangle"
        elif len(approx) == 4:
            shape_type = "rectangle"
        elif len(approx) > 4:
            area = cv2.contourArea(contour)
            perimeter = cv2.arcLength(contour, True)
            if perimeter > 0:
                circularity = 4 * np.pi * area / (perimeter ** 2)
                shape_type = "circle" if circularity > 0.8 else "polygon"
        return shape_type

    @staticmethod
    def _device_contours(image: np.ndarray):
        """image -> contours with cv2.contourArea >= 100, as cv2.findContours would list them (src L76-78 / L140-142)."""
        img = _runtime.as_bgr_u8(image)
        with _runtime.lock():
            headers, points = _contours.shape_contours_host(_runtime.context(), img, ShapeAnalyzer.CANNY_LOW,
                                                            ShapeAnalyzer.CANNY_HIGH)
        return _contours.to_cv2_contours(headers, points)

    def extract_shapes(self, image: np.ndarray) -> List[Dict[str, Any]]:
        """src L63-123: [{'type', 'coordinates': [{'x','y'}, ...]}]."""
        shape_results = []
        for contour in self._device_contours(image):
            shape_type = self._classify(contour)
            coordinates = [{"x": int(point[0][0]), "y": int(point[0][1])} for point in contour]
            shape_results.append({"type": shape_type, "coordinates": coordinates})
        return shape_results

    @staticmethod
    def analyze_shapes(image: np.ndarray) -> Dict[str, Any]:
        """src L125-189: {'shapes': [...], 'total_shapes', 'metadata': {'image_width','image_height'}}."""
        return ShapeAnalyzer.shapes_from_contours(ShapeAnalyzer._device_contours(image), image.shape[1], image.shape[0])

    @staticmethod
    def shapes_from_mask(preprocessed: np.ndarray, image_width: int, image_height: int) -> Dict[str, Any]:
        """The tail of analyze_shapes (src L134-189) on an already computed edge mask in host memory."""
        with _runtime.lock():
            headers, points = _contours.find_external_host(_runtime.context(), preprocessed, _contours.REFERENCE_MIN_AREA2)
        return ShapeAnalyzer.shapes_from_contours(_contours.to_cv2_contours(headers, points), image_width, image_height)

    @staticmethod
    def shapes_from_contours(contours, image_width: int, image_height: int) -> Dict[str, Any]:
        """src L142-189 on the contours that passed `cv2.contourArea(contour) < 100: continue`."""
        shape_results = []
        for contour in contours:
            x, y, w, h = cv2.boundingRect(contour)
            border_radius = ShapeAnalyzer.detect_border_radius(contour)
            shape_type = ShapeAnalyzer._classify(contour)
            shape_results.append({"type": shape_type, "x": x, "y": y, "width": w, "height": h,
                                  "border_radius": border_radius, "area": cv2.contourArea(contour)})
        return {"shapes": shape_results, "total_shapes": len(shape_results),
                "metadata": {"image_width": image_width, "image_height": image_height}}
