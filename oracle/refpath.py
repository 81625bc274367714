"""The reference's call sequences for the hot path, written directly against
the library the reference calls (cv2 / numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the "port" arm: same
third-party binary (opencv-python-headless 4.13.0.92), same call order and
arguments as the reference's services, so it can stand in for the reference
on the GPU box where /root/reference does not exist.  It is used (a) as the
CPU baseline timed by bench.py and (b) as an independent check of
oracle/cvops.py.  Citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import cv2

KMEANS_CRITERIA = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 200, 0.2)


def shape_mask(bgr: np.ndarray) -> np.ndarray:
    """ShapeAnalyzer.preprocess_image -- shape_analyzer.pyc src L18-28."""
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    blurred = cv2.GaussianBlur(gray, (5, 5), 0)
    edges = cv2.Canny(blurred, 50, 150)
    return cv2.dilate(edges, np.ones((3, 3), np.uint8), iterations=1)


def shadow_level(bgr: np.ndarray) -> str:
    """ShadowAnalyzer.analyze_shadow_level -- shadow_analyzer.pyc src L8-31."""
    processed = cv2.GaussianBlur(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY), (5, 5), 0)
    thresh = cv2.adaptiveThreshold(processed, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, 2)
    shadow_pixels = processed[thresh == 255]
    if shadow_pixels.size == 0:
        return "Low"
    avg_darkness = 255 - np.mean(shadow_pixels)
    if avg_darkness < 30:
        return "Low"
    if avg_darkness < 60:
        return "Moderate"
    return "High"


def shadow_mask(bgr: np.ndarray):
    processed = cv2.GaussianBlur(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY), (5, 5), 0)
    thresh = cv2.adaptiveThreshold(processed, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, 2)
    return processed, thresh


def text_mask(bgr: np.ndarray) -> np.ndarray:
    """TextExtractor.preprocess_image -- text_extractor.py:26-46."""
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY) if (bgr.ndim > 2 and bgr.shape[2] > 1) else bgr
    height, width = gray.shape[:2]
    if height < 30 or width < 100:
        scale = max(2, 300 / width, 100 / height)
        gray = cv2.resize(gray, None, fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    _, binary = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    if np.mean(binary) > 127:
        binary = cv2.bitwise_not(binary)
    return binary


def font_mask(bgr: np.ndarray) -> np.ndarray:
    """FontDetector.preprocess_image -- font_detector.py:28-35."""
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    return cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, 2)


def auto_resize(bgr: np.ndarray) -> np.ndarray:
    """`auto` branch of validate_and_preprocess_image -- utils.py:120-127."""
    h, w = bgr.shape[:2]
    if max(h, w) > 2000:
        scale = 2000 / max(h, w)
        bgr = cv2.resize(bgr, (int(w * scale), int(h * scale)), interpolation=cv2.INTER_AREA)
    return bgr


def dominant_colors(pixels: np.ndarray, n_colors: int):
    """ColorExtractor._get_dominant_colors -- color_extractor.py:177-197."""
    unique_colors = np.unique(pixels, axis=0)
    k = min(n_colors, len(unique_colors))
    if k <= 1:
        return unique_colors, np.array([0] * len(unique_colors))
    _, labels, centers = cv2.kmeans(np.float32(unique_colors), k, None, KMEANS_CRITERIA, 10, cv2.KMEANS_PP_CENTERS)
    return centers.astype(np.uint8), labels.flatten()


def extract_colors(bgr: np.ndarray, n_colors: int = 5) -> dict:
    """ColorExtractor.extract_colors -- color_extractor.py:217-284 (ndarray
    input with H>4).  Uses the process-global np.random / cv2 RNG like the
    reference: seed with np.random.seed(s); cv2.setRNGSeed(s)."""
    from .cvops import palette_tail

    img = cv2.cvtColor(bgr.copy(), cv2.COLOR_BGR2RGB)
    pixels = img.reshape(-1, 3)
    noise = np.random.normal(0, 0.5, pixels.shape).astype(np.int8)
    pixels = np.clip(pixels.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    centers, labels = dominant_colors(pixels, n_colors)
    return palette_tail(centers, labels)


def kmeans_pixels(pixels_f32: np.ndarray, init_centers: np.ndarray, max_iter: int = 200, eps: float = 0.2):
    """Per-pixel Lloyd from given centres through cv2 itself (SURVEY.md A.8
    'seeded mode'): labels0 = nearest(init) + KMEANS_USE_INITIAL_LABELS."""
    from .cvops import assign

    lab0, _ = assign(pixels_f32, init_centers.astype(np.float32))
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, max_iter, eps)
    comp, labels, centers = cv2.kmeans(pixels_f32, len(init_centers), lab0.reshape(-1, 1).copy(), crit, 1,
                                       cv2.KMEANS_USE_INITIAL_LABELS)
    return centers, labels.ravel(), comp


def analyze_shapes_from_mask(preprocessed: np.ndarray, image_shape) -> dict:
    """Host tail of ShapeAnalyzer.analyze_shapes -- shape_analyzer.pyc src
    L140-189: contours + polygon geometry on the dilated mask."""
    contours, _ = cv2.findContours(preprocessed, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    shapes = []
    for contour in contours:
        if cv2.contourArea(contour) < 100:
            continue
        x, y, w, h = cv2.boundingRect(contour)
        eps_r = 0.02 * cv2.arcLength(contour, True)
        approx_r = cv2.approxPolyDP(contour, eps_r, True)
        border_radius = 0.0
        if len(approx_r) > 4:
            hull_area = cv2.contourArea(cv2.convexHull(contour))
            contour_area = cv2.contourArea(contour)
            if hull_area > 0:
                border_radius = max(0.0, (1 - contour_area / hull_area) * 50.0)
        approx = cv2.approxPolyDP(contour, 0.04 * cv2.arcLength(contour, True), True)
        shape_type = "unknown"
        if len(approx) == 3:
            shape_type = "triangle"
        elif len(approx) == 4:
            shape_type = "rectangle"
        elif len(approx) > 4:
            area = cv2.contourArea(contour)
            perimeter = cv2.arcLength(contour, True)
            if perimeter > 0:
                shape_type = "circle" if 4 * np.pi * area / perimeter ** 2 > 0.8 else "polygon"
        shapes.append({"type": shape_type, "x": x, "y": y, "width": w, "height": h,
                       "border_radius": border_radius, "area": cv2.contourArea(contour)})
    return {"shapes": shapes, "total_shapes": len(shapes),
            "metadata": {"image_width": image_shape[1], "image_height": image_shape[0]}}
