"""Process-wide llfe context for the service classes (no torch needed).

The reference handles one request at a time per worker process; likewise all
service calls of a process share one context (device LLFE_DEVICE, default 0)
and are serialised by a lock.
"""
from __future__ import annotations

import os
import threading

import numpy as np

from .._native import Context

_ctx: Context | None = None
_lock = threading.RLock()


def context() -> Context:
    global _ctx
    with _lock:
        if _ctx is None:
            _ctx = Context(int(os.environ.get("LLFE_DEVICE", "0")))
        return _ctx


def lock() -> threading.RLock:
    return _lock


def as_bgr_u8(image: np.ndarray) -> np.ndarray:
    """C-contiguous (H, W, 3) uint8 BGR array for the kernels, with the input contract of the cv2 calls the
    service methods replace: `cv2.cvtColor(image, COLOR_BGR2GRAY)` followed by an 8-bit-only op (Canny,
    adaptiveThreshold, Otsu).  3-channel uint8 passes through; 4-channel uint8 (BGRA) is accepted and its alpha
    ignored, as cvtColor does; everything else -- not an ndarray, another channel count, another depth, an empty
    image -- raises `cv2.error`, the exception the reference's callers see."""
    import cv2

    if not isinstance(image, np.ndarray):
        raise cv2.error("cvtColor: Overload resolution failed: src is not a numpy array, neither a scalar")
    if image.ndim != 3 or image.shape[2] not in (3, 4):
        raise cv2.error(f"cvtColor: Invalid number of channels in input image: 'VScn::contains(scn)', shape {image.shape}")
    if image.shape[0] == 0 or image.shape[1] == 0:
        raise cv2.error("cvtColor: (-215:Assertion failed) !_src.empty()")
    if image.dtype != np.uint8:
        raise cv2.error(f"(-215:Assertion failed) depth == CV_8U: unsupported image depth {image.dtype}")
    if image.shape[2] == 4:
        image = image[:, :, :3]
    return np.ascontiguousarray(image)
