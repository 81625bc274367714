"""Synthetic "design-style" inputs.

These are the images every benchmark and most parity tests run on: light
background, gradient header, flat rectangles, anti-aliased circles, dark
"text" bars and +-2 sensor noise, as u8 BGR HWC (the OpenCV convention the
reference's services take, e.g. app/services/analyze/color_extractor.py:204).
Only numpy + cv2 drawing calls; no reference code involved.
"""
from __future__ import annotations

import numpy as np


def design_image(h: int = 1080, w: int = 1920, seed: int = 0, noise: bool = True) -> np.ndarray:
    """One (h, w, 3) u8 BGR design-style image, deterministic in `seed`."""
    import cv2

    r = np.random.default_rng(seed)
    img = np.empty((h, w, 3), np.uint8)
    img[:] = r.integers(225, 250, 3)
    gx = np.linspace(0, 1, w)[None, :, None]
    c0 = r.integers(0, 255, 3)
    c1 = r.integers(0, 255, 3)
    img[: h // 8] = (c0 * (1 - gx) + c1 * gx).astype(np.uint8)
    for _ in range(12):
        x0, y0 = int(r.integers(0, max(1, w - 50))), int(r.integers(h // 8, max(h // 8 + 1, h - 50)))
        ww, hh = int(r.integers(40, max(41, w // 3))), int(r.integers(30, max(31, h // 3)))
        col = tuple(int(v) for v in r.integers(0, 255, 3))
        cv2.rectangle(img, (x0, y0), (min(x0 + ww, w - 1), min(y0 + hh, h - 1)), col, -1)
    for _ in range(6):
        col = tuple(int(v) for v in r.integers(0, 255, 3))
        cv2.circle(img, (int(r.integers(0, w)), int(r.integers(0, h))),
                   int(r.integers(15, max(16, h // 6))), col, -1, lineType=cv2.LINE_AA)
    for _ in range(30):
        y = int(r.integers(h // 8, max(h // 8 + 1, h - 10)))
        x = int(r.integers(0, max(1, w - 300)))
        cv2.rectangle(img, (x, y), (x + int(r.integers(50, 300)), y + int(r.integers(3, 9))), (40, 40, 40), -1)
    if noise:
        img = np.clip(img.astype(np.int16) + r.integers(-2, 3, img.shape), 0, 255).astype(np.uint8)
    return img


def design_batch(n: int, h: int = 1080, w: int = 1920, seed0: int = 0, distinct: int | None = None) -> np.ndarray:
    """(n, h, w, 3) u8 batch; image i uses seed `seed0 + (i % distinct)`.

    `distinct` bounds host generation time for large synthetic batches: the
    remaining images are copies of the first `distinct` ones rolled by a few
    rows so that no two images in the batch are byte-identical.
    """
    distinct = n if distinct is None else max(1, min(distinct, n))
    base = [design_image(h, w, seed0 + i) for i in range(distinct)]
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        src = base[i % distinct]
        k = i // distinct
        out[i] = src if k == 0 else np.roll(src, 7 * k, axis=0)
    return out


def noise_image(h: int, w: int, seed: int = 0) -> np.ndarray:
    """Adversarial input: uniform random u8 BGR."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
