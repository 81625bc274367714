"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the Pillow calls on the reference's
`ImageProcessor.auto_process_image` path (/root/reference/app/services/analyze/image_processor.py:205-225):

    pil_image = Image.fromarray(cv2.cvtColor(image, cv2.COLOR_BGR2RGB))
    pil_image.thumbnail((max_width, max_height), Image.Resampling.LANCZOS)      # reducing_gap = 2.0
    return cv2.cvtColor(np.array(pil_image), cv2.COLOR_RGB2BGR)

Pillow (pinned `pillow<11` in the reference's requirements.txt:11; 12.2.0 installed here) is a third-party dependency
whose C sources (src/libImaging/Resample.c, Reduce.c) are not in the container; the arithmetic below restates the
published algorithm and is pinned against the installed binary by tests/test_oracle_pil.py:

  * thumbnail size: `Image.thumbnail`'s `preserve_aspect_ratio` (PIL/Image.py), literally;
  * `Image.resize(size, LANCZOS, box, reducing_gap=2.0)`: an integer `reduce` by `int(scale / 2)` per axis over the
    "safe box" when that is > 1, then the two-pass resample;
  * `ImagingReduce`: out = ((sum + n // 2) * floor(2^24 / n)) >> 24 with n = pixels under the cell, partial cells at
    the right / bottom edge with their own n (rule determined by probing the installed binary);
  * `ImagingResample` 8 bits per channel: per output index the window [xmin, xmin + xmax) and double-precision weights
    lanczos((x + xmin - center + 0.5) / filterscale) normalised to sum 1, converted to 22-bit fixed point with
    round-half-away-from-zero; horizontal pass first (over the rows the vertical pass needs), clipped to u8, then the
    vertical pass; out = clip8((2^21 + sum(pixel * k)) >> 22).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def thumbnail_size(width: int, height: int, size) -> tuple[int, int] | None:
    """PIL.Image.Image.thumbnail's preserve_aspect_ratio(): final (w, h), or None when the image already fits."""
    x, y = (math.floor(size[0]), math.floor(size[1]))

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    if x >= width and y >= height:
        return None
    aspect = width / height
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return x, y


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """-> (ksize, bounds (out_size, 2) int [xmin, count], kk (out_size, ksize) int32 fixed-point weights).
    in0 / in1 are C floats in Pillow's signature (`float in0, float in1`)."""
    in0 = float(np.float32(in0))
    in1 = float(np.float32(in1))
    scale = float(np.float32(in1) - np.float32(in0)) / out_size      # (double)(in1 - in0) / outSize, float subtraction
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_lanczos(img: np.ndarray, out_w: int, out_h: int, box=None) -> np.ndarray:
    """ImagingResample(im, out_w, out_h, LANCZOS, box) for (H, W, C) or (H, W) uint8."""
    squeeze = img.ndim == 2
    src = img[:, :, None] if squeeze else img
    h, w, _ = src.shape
    if box is None:
        box = (0, 0, w, h)
    need_h = out_w != w or box[0] != 0 or box[2] != out_w
    need_v = out_h != h or box[1] != 0 or box[3] != out_h
    _, bh, kh = precompute_coeffs(w, box[0], box[2], out_w)
    _, bv, kv = precompute_coeffs(h, box[1], box[3], out_h)
    first = int(bv[0, 0])
    last = int(bv[out_h - 1, 0] + bv[out_h - 1, 1])
    cur = src
    if need_h:
        rows = cur[first:last].astype(np.int64)
        tmp = np.empty((last - first, out_w, src.shape[2]), np.uint8)
        for xx in range(out_w):
            x0, n = int(bh[xx, 0]), int(bh[xx, 1])
            acc = (rows[:, x0:x0 + n, :] * kh[xx, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
        bv = bv.copy()
        bv[:, 0] -= first
    if need_v:
        rows = cur.astype(np.int64)
        out = np.empty((out_h, cur.shape[1], src.shape[2]), np.uint8)
        for yy in range(out_h):
            y0, n = int(bv[yy, 0]), int(bv[yy, 1])
            acc = (rows[y0:y0 + n] * kv[yy, :n].astype(np.int64)[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        cur = out
    cur = np.ascontiguousarray(cur)
    return cur[:, :, 0] if squeeze else cur


def reduce_box_mean(img: np.ndarray, fx: int, fy: int, box=None) -> np.ndarray:
    """ImagingReduce(im, fx, fy, box): integer box means, out = ((sum + n // 2) * floor(2^24 / n)) >> 24 with n the number
    of pixels under the (possibly partial, at the right / bottom edge of the box) fx x fy cell."""
    squeeze = img.ndim == 2
    src = img[:, :, None] if squeeze else img
    if box is not None:
        src = src[box[1]:box[3], box[0]:box[2]]
    h, w, c = src.shape
    oh, ow = (h + fy - 1) // fy, (w + fx - 1) // fx
    pad = np.zeros((oh * fy, ow * fx, c), np.int64)
    pad[:h, :w] = src
    sums = pad.reshape(oh, fy, ow, fx, c).sum((1, 3))
    ny = np.minimum(fy, h - np.arange(oh) * fy)
    nx = np.minimum(fx, w - np.arange(ow) * fx)
    n = (ny[:, None] * nx[None, :])[:, :, None]
    out = (((sums + n // 2) * ((1 << 24) // n)) >> 24).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def safe_box(size, box, im_size):
    """PIL.Image.Image._get_safe_box for LANCZOS."""
    fs = LANCZOS_SUPPORT - 0.5
    sx = (box[2] - box[0]) / size[0]
    sy = (box[3] - box[1]) / size[1]
    return (max(0, int(box[0] - fs * sx)), max(0, int(box[1] - fs * sy)),
            min(im_size[0], math.ceil(box[2] + fs * sx)), min(im_size[1], math.ceil(box[3] + fs * sy)))


def resize_lanczos(img: np.ndarray, size, box=None, reducing_gap=2.0) -> np.ndarray:
    """PIL.Image.Image.resize(size, LANCZOS, box, reducing_gap) for "L" / "RGB" arrays (PIL/Image.py)."""
    h, w = img.shape[:2]
    if box is None:
        box = (0, 0, w, h)
    size = tuple(size)
    if (w, h) == size and tuple(box) == (0, 0, w, h):
        return img.copy()
    if reducing_gap is not None:
        fx = int((box[2] - box[0]) / size[0] / reducing_gap) or 1
        fy = int((box[3] - box[1]) / size[1] / reducing_gap) or 1
        if fx > 1 or fy > 1:
            rb = safe_box(size, box, (w, h))
            img = reduce_box_mean(img, fx, fy, rb)
            box = ((box[0] - rb[0]) / fx, (box[1] - rb[1]) / fy, (box[2] - rb[0]) / fx, (box[3] - rb[1]) / fy)
            h, w = img.shape[:2]
    if h > w * 100 and size[1] < h:
        img = resample_lanczos(img, w, size[1], (0, box[1], w, box[3]))
        return resample_lanczos(img, size[0], size[1], (box[0], 0, box[2], size[1]))
    return resample_lanczos(img, size[0], size[1], box)


def thumbnail_lanczos(img: np.ndarray, max_width: int, max_height: int) -> np.ndarray:
    """`im.thumbnail((max_width, max_height), Image.Resampling.LANCZOS)` on an (H, W, 3) / (H, W) uint8 array
    (channel order is irrelevant: every channel is filtered alike)."""
    h, w = img.shape[:2]
    final = thumbnail_size(w, h, (max_width, max_height))
    if final is None or final == (w, h):
        return img.copy()
    return resize_lanczos(img, final, None, 2.0)
