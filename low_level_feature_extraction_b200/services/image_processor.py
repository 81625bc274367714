"""Drop-in for the hot-path part of the reference's ImageProcessor
(app/services/analyze/image_processor.py): `load_cv2_image` (:42-81, host decode)
and `resize_image` (:83-118, INTER_AREA on the GPU).  WebP / format conversion
(:121-274) is codec work outside the path."""
from __future__ import annotations

import logging
import time
from typing import Union

import cv2
import numpy as np
from PIL import Image

from . import _runtime

logger = logging.getLogger(__name__)


def resize_area(image: np.ndarray, new_width: int, new_height: int) -> np.ndarray:
    """cv2.resize(image, (new_width, new_height), interpolation=cv2.INTER_AREA) for down-scaling u8 images."""
    if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] not in (1, 3)):
        raise TypeError("Unsupported image type")
    src = np.ascontiguousarray(image)
    sh, sw = src.shape[:2]
    c = 1 if src.ndim == 2 else src.shape[2]
    if new_width <= 0 or new_height <= 0:
        raise ValueError("Invalid target size")
    if new_width > sw or new_height > sh:
        raise NotImplementedError("INTER_AREA up-scaling is outside the hot path")
    dst = np.empty((new_height, new_width) + src.shape[2:], np.uint8)
    with _runtime.lock():
        _runtime.context().call("llfe_resize_area_host", src, sh, sw, c, dst, new_height, new_width)
    return dst


def resize_linear(image: np.ndarray, new_width: int, new_height: int) -> np.ndarray:
    """cv2.resize(image, (new_width, new_height), interpolation=cv2.INTER_LINEAR) for u8 images."""
    if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] not in (1, 3)):
        raise TypeError("Unsupported image type")
    if new_width <= 0 or new_height <= 0:
        raise ValueError("Invalid target size")
    src = np.ascontiguousarray(image)
    sh, sw = src.shape[:2]
    c = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((new_height, new_width) + src.shape[2:], np.uint8)
    with _runtime.lock():
        _runtime.context().call("llfe_resize_linear_host", src, sh, sw, c, dst, new_height, new_width)
    return dst


def resize_lanczos4(image: np.ndarray, new_width: int, new_height: int) -> np.ndarray:
    """cv2.resize(image, (new_width, new_height), interpolation=cv2.INTER_LANCZOS4) for u8 images."""
    if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] not in (1, 3)):
        raise TypeError("Unsupported image type")
    if new_width <= 0 or new_height <= 0:
        raise ValueError("Invalid target size")
    src = np.ascontiguousarray(image)
    sh, sw = src.shape[:2]
    c = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((new_height, new_width) + src.shape[2:], np.uint8)
    with _runtime.lock():
        _runtime.context().call("llfe_resize_lanczos4_host", src, sh, sw, c, dst, new_height, new_width)
    return dst


class ImageProcessor:
    logger = logging.getLogger(__name__)

    @staticmethod
    def load_cv2_image(image_bytes: bytes) -> np.ndarray:
        """bytes -> BGR ndarray; ValueError on any failure (image_processor.py:42-81)."""
        start_time = time.time()
        try:
            if image_bytes is None or len(image_bytes) == 0:
                raise ValueError("No image bytes provided")
            image = cv2.imdecode(np.frombuffer(image_bytes, np.uint8), cv2.IMREAD_COLOR)
            if image is None:
                raise ValueError("Failed to decode image")
            return image
        except Exception as e:
            ImageProcessor.logger.error(f"Image loading error: {e}")
            raise ValueError("Decoded image is empty")
        finally:
            ImageProcessor.logger.info(f"Image loading took {time.time() - start_time:.4f} seconds")

    @staticmethod
    def resize_image(image: Union[Image.Image, np.ndarray], max_width: int = 1920, max_height: int = 1080,
                     fit: str = "max") -> Union[Image.Image, np.ndarray]:
        """Aspect-preserving fit into (max_width, max_height), never up-scaling (image_processor.py:83-118)."""
        if isinstance(image, Image.Image):
            image.thumbnail((max_width, max_height))
            return image
        if isinstance(image, np.ndarray):
            height, width = image.shape[:2]
            scaling_factor = min(max_width / width, max_height / height, 1)
            new_width = int(width * scaling_factor)
            new_height = int(height * scaling_factor)
            return resize_area(image, new_width, new_height)
        raise TypeError("Unsupported image type")
