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
    """C-contiguous (H, W, 3) uint8 view/copy of a BGR image; raises like cv2 on bad input."""
    if not isinstance(image, np.ndarray):
        raise TypeError("Expected a numpy.ndarray image")
    if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
        raise ValueError(f"Expected an (H, W, 3) uint8 BGR image, got shape {image.shape} dtype {image.dtype}")
    if image.shape[0] == 0 or image.shape[1] == 0:
        raise ValueError("Empty image")
    return np.ascontiguousarray(image)
