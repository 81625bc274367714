"""Host side of the device contour finder (csrc/k_contours.cu, include/llfe.h `llfe_contours_external*`):
buffer sizing + retry, and the conversion of header / point records into what `cv2.findContours(mask,
RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)` returns (list of (N, 1, 2) int32 arrays, in cv2's order)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

HEADER = np.dtype([("start", "<i4"), ("npts", "<i4"), ("offset", "<i4"), ("minx", "<i4"), ("miny", "<i4"),
                   ("maxx", "<i4"), ("maxy", "<i4"), ("pad", "<i4"), ("area2", "<i8")])
assert HEADER.itemsize == 40

# the reference keeps contours with cv2.contourArea >= 100 (shape_analyzer pyc L78 / L142); area2 = 2 * area
REFERENCE_MIN_AREA2 = 200
NO_POINTS = 1 << 62   # min_area2 that no contour reaches: headers only


def run_with_retry(call: Callable[[np.ndarray, int, Optional[np.ndarray], int, np.ndarray], None], want_points: bool,
                   max_contours: int = 4096, max_points: int = 1 << 16) -> Tuple[np.ndarray, np.ndarray]:
    """call(headers, max_contours, points | None, max_points, counts) runs the C entry point; it is repeated with
    larger buffers when the device reports more contours / points than fit.  -> (headers[:n], points)."""
    while True:
        headers = np.empty(max_contours, HEADER)
        points = np.empty((max_points, 2), np.int32) if want_points else None
        counts = np.zeros(4, np.int32)
        call(headers, max_contours, points, max_points if want_points else 0, counts)
        if counts[0] > max_contours or counts[2]:
            max_contours = max(max_contours, int(counts[0]))
            if counts[2]:
                max_points = max(4 * max_points, int(counts[1]))
            continue
        return headers[:counts[0]], points


def in_cv2_order(headers: np.ndarray) -> np.ndarray:
    """cv2 lists the contours in reverse order of discovery = descending first pixel."""
    return headers[np.argsort(-headers["start"].astype(np.int64), kind="stable")]


def to_cv2_contours(headers: np.ndarray, points: np.ndarray) -> List[np.ndarray]:
    """Contours whose points were written (those that passed the area threshold), in cv2's order and layout."""
    return [points[h["offset"]:h["offset"] + h["npts"]].reshape(-1, 1, 2).copy()
            for h in in_cv2_order(headers) if h["offset"] >= 0]


def bounding_rects(headers: np.ndarray) -> List[Tuple[int, int, int, int]]:
    """cv2.boundingRect of every external contour, in cv2's contour order."""
    hs = in_cv2_order(headers)
    return [(int(h["minx"]), int(h["miny"]), int(h["maxx"] - h["minx"] + 1), int(h["maxy"] - h["miny"] + 1)) for h in hs]


def find_external_host(ctx, mask: np.ndarray, min_area2: int = 0, want_points: bool = True):
    """Device contours of a host mask (any non-zero byte is foreground).  -> (headers, points)."""
    mask = np.ascontiguousarray(mask)
    assert mask.ndim == 2 and mask.dtype == np.uint8
    h, w = mask.shape

    def call(headers, mc, points, mp, counts):
        ctx.call("llfe_contours_external_host", mask, h, w, int(min_area2), headers, mc, points, mp, counts)

    return run_with_retry(call, want_points)


def shape_contours_host(ctx, bgr: np.ndarray, low: int, high: int, min_area2: int = REFERENCE_MIN_AREA2,
                        mask_out: Optional[np.ndarray] = None):
    """ShapeAnalyzer's whole device part in one call: BGR image -> dilated Canny mask -> external contours."""
    h, w = bgr.shape[:2]

    def call(headers, mc, points, mp, counts):
        ctx.call("llfe_shape_contours_host", bgr, h, w, low, high, int(min_area2), mask_out, headers, mc, points, mp, counts)

    return run_with_retry(call, True)
