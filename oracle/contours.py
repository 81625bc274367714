"""TEST INFRASTRUCTURE ONLY -- CPU restatement of `cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)`.

The reference calls it in ShapeAnalyzer.extract_shapes / analyze_shapes (app/services/__pycache__/
shape_analyzer.cpython-312.pyc, source lines 76 and 140) and in FontDetector.detect_text_regions
(app/services/analyze/font_detector.py:51-55).  The algorithm lives in OpenCV (PyPI opencv-python-headless,
requirements.txt:10; installed 4.13.0.92), which is not vendored: this file restates its published raster scan +
Suzuki/Abe border following (modules/imgproc/src/contours.cpp: cvFindNextContour, icvFetchContour) and is pinned
against the installed binary in tests/test_contours_oracle.py (thousands of random masks, point for point).

`find_external` is the literal scan (marks borders in the image like OpenCV does); `external_starts_ideal` is the
closed form the CUDA path uses for WHICH components are external and WHERE their borders start (raster-first pixel
of every 8-connected component whose left neighbour belongs to the frame-connected 4-connected background);
the same test file proves both equal to cv2 on the same masks.
"""
from __future__ import annotations

import numpy as np

# chain code: 0 = E, 1 = NE, 2 = N, 3 = NW, 4 = W, 5 = SW, 6 = S, 7 = SE (y grows downwards)
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def _trace(img, x0, y0, nbd):
    """icvFetchContour on the 1-pixel padded int image; returns CHAIN_APPROX_SIMPLE points in unpadded coordinates."""
    pts = []
    s_end = s = 4
    while True:
        s = (s - 1) & 7
        if img[y0 + DY[s], x0 + DX[s]] != 0 or s == s_end:
            break
    if s == s_end and img[y0 + DY[s], x0 + DX[s]] == 0:   # single pixel
        img[y0, x0] = -nbd
        return [(x0 - 1, y0 - 1)]
    i1 = (x0 + DX[s], y0 + DY[s])
    i3 = (x0, y0)
    prev_s = s ^ 4
    x, y = x0, y0
    while True:
        s_end = s
        while True:
            s += 1
            i4 = (i3[0] + DX[s & 7], i3[1] + DY[s & 7])
            if img[i4[1], i4[0]] != 0:
                break
        s &= 7
        if ((s - 1) & 0xffffffff) < s_end:      # the search passed the east neighbour and found it empty
            img[i3[1], i3[0]] = -nbd
        elif img[i3[1], i3[0]] == 1:
            img[i3[1], i3[0]] = nbd
        if s != prev_s:
            pts.append((x - 1, y - 1))
        prev_s = s
        x += DX[s]
        y += DY[s]
        if i4 == (x0, y0) and i3 == i1:
            break
        i3 = i4
        s = (s + 4) & 7
    return pts


def find_external(mask: np.ndarray):
    """-> list of point lists [(x, y), ...] in cv2's order (last found first)."""
    h, w = mask.shape
    img = np.zeros((h + 2, w + 2), np.int64)
    img[1:-1, 1:-1] = mask != 0
    nbd = 1
    out = []
    for y in range(1, h + 1):
        prev = 0
        lnbd_x = 0
        for x in range(1, w + 1):
            p = int(img[y, x])
            if p == prev:
                continue
            ok = True
            is_hole = False
            if not (prev == 0 and p == 1):
                if p != 0 or prev < 1:
                    ok = False
                else:
                    if prev & ~1:
                        lnbd_x = x - 1
                    is_hole = True
            if ok and (is_hole or img[y, lnbd_x] > 0):   # RETR_EXTERNAL: no holes, nothing inside a traced border
                ok = False
            if ok:
                nbd += 1
                out.append(_trace(img, x, y, nbd))
                p = int(img[y, x])
            prev = p
            if prev & ~1:
                lnbd_x = x
    return out[::-1]


def external_starts_ideal(mask: np.ndarray):
    """Start pixels (x, y) of the external contours by the closed form, ascending raster order."""
    from scipy import ndimage as ndi

    fg = mask != 0
    h, w = fg.shape
    pad = np.zeros((h + 2, w + 2), bool)
    pad[1:-1, 1:-1] = fg
    bl, _ = ndi.label(~pad)                            # 4-connected background
    outer = bl == bl[0, 0]
    fl, nf = ndi.label(pad, structure=np.ones((3, 3)))  # 8-connected foreground
    first = ndi.minimum(np.arange(pad.size).reshape(pad.shape), fl, index=np.arange(1, nf + 1)) if nf else []
    starts = []
    for lin in np.atleast_1d(first):
        y, x = divmod(int(lin), w + 2)
        if outer[y, x - 1]:
            starts.append((x - 1, y - 1))
    return sorted(starts, key=lambda p: (p[1], p[0]))


def contour_area2(points) -> int:
    """2 * signed area by Green's formula (|.| / 2 == cv2.contourArea for integer points)."""
    a = 0
    n = len(points)
    for i in range(n):
        x0, y0 = points[i - 1]
        x1, y1 = points[i]
        a += x0 * y1 - y0 * x1
    return a
