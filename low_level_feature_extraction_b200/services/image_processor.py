"""Drop-in for the hot-path part of the reference's ImageProcessor
(app/services/analyze/image_processor.py): `load_cv2_image` (:42-81; PNG reconstruction on the
GPU, services/png.py), `resize_image` (:83-118, INTER_AREA on the GPU) and `auto_process_image`
(:183-229; decode + Pillow's LANCZOS thumbnail on the GPU).  WebP / format conversion
(:121-181, :231-274) is codec work outside the path."""
from __future__ import annotations

import logging
import math
import time
from typing import Union

import numpy as np
from PIL import Image

from . import _runtime
from .png import imdecode_color

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


# ---- Pillow's thumbnail(size, LANCZOS): PIL/Image.py's Python layer around the two device primitives -----------------------
_LANCZOS_SUPPORT = 3.0


def _thumbnail_size(width: int, height: int, size) -> tuple[int, int] | None:
    """Image.thumbnail's preserve_aspect_ratio(): the final (width, height), None when the image already fits."""
    x, y = math.floor(size[0]), math.floor(size[1])

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


def _pil_resize_call(src: np.ndarray, fx: int, fy: int, reduce_box, box, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    c = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((dh, dw) + src.shape[2:], np.uint8)
    rb = np.asarray(reduce_box if reduce_box is not None else (0, 0, sw, sh), np.int32)
    fb = np.asarray(box, np.float32)
    with _runtime.lock():
        _runtime.context().call("llfe_pil_resize_lanczos_host", src, sh, sw, c, int(fx), int(fy), rb, fb, dst, int(dh), int(dw))
    return dst


def pil_resize_lanczos(image: np.ndarray, size, box=None, reducing_gap: float | None = 2.0) -> np.ndarray:
    """np.array(Image.fromarray(image).resize(size, Image.Resampling.LANCZOS, box, reducing_gap)) for "L" / "RGB" uint8
    arrays: Image.resize's factor / box rules (PIL/Image.py), the integer reduction and the resampling on the GPU."""
    if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] != 3):
        raise TypeError("Unsupported image type")
    if reducing_gap is not None and reducing_gap < 1.0:
        raise ValueError("reducing_gap must be 1.0 or greater")
    src = np.ascontiguousarray(image)
    h, w = src.shape[:2]
    size = (int(size[0]), int(size[1]))
    if size[0] <= 0 or size[1] <= 0:
        raise ValueError("height and width must be > 0")
    if box is None:
        box = (0, 0, w, h)
    if (w, h) == size and tuple(box) == (0, 0, w, h):
        return src.copy()
    fx = fy = 1
    reduce_box = None
    if reducing_gap is not None:
        fx = int((box[2] - box[0]) / size[0] / reducing_gap) or 1
        fy = int((box[3] - box[1]) / size[1] / reducing_gap) or 1
        if fx > 1 or fy > 1:
            fs = _LANCZOS_SUPPORT - 0.5          # Image._get_safe_box
            sx = (box[2] - box[0]) / size[0]
            sy = (box[3] - box[1]) / size[1]
            reduce_box = (max(0, int(box[0] - fs * sx)), max(0, int(box[1] - fs * sy)),
                          min(w, math.ceil(box[2] + fs * sx)), min(h, math.ceil(box[3] + fs * sy)))
            box = ((box[0] - reduce_box[0]) / fx, (box[1] - reduce_box[1]) / fy,
                   (box[2] - reduce_box[0]) / fx, (box[3] - reduce_box[1]) / fy)
            w = -(-(reduce_box[2] - reduce_box[0]) // fx)
            h = -(-(reduce_box[3] - reduce_box[1]) // fy)
    if h > w * 100 and size[1] < h:
        # Image.resize's two-step path for very tall images: rows first, then columns
        mid = _pil_resize_call(src, fx, fy, reduce_box, (0, box[1], w, box[3]), w, size[1])
        return _pil_resize_call(mid, 1, 1, None, (box[0], 0, box[2], size[1]), size[0], size[1])
    return _pil_resize_call(src, fx, fy, reduce_box, box, size[0], size[1])


def pil_thumbnail_lanczos(image: np.ndarray, max_width: int, max_height: int) -> np.ndarray:
    """`im = Image.fromarray(image); im.thumbnail((max_width, max_height), Image.Resampling.LANCZOS); np.array(im)`."""
    h, w = image.shape[:2]
    final = _thumbnail_size(w, h, (max_width, max_height))
    if final is None or final == (w, h):
        return np.ascontiguousarray(image).copy()
    return pil_resize_lanczos(image, final, None, 2.0)


class ImageProcessor:
    logger = logging.getLogger(__name__)

    @staticmethod
    def load_cv2_image(image_bytes: bytes) -> np.ndarray:
        """bytes -> BGR ndarray; ValueError on any failure (image_processor.py:42-81)."""
        start_time = time.time()
        try:
            if image_bytes is None or len(image_bytes) == 0:
                raise ValueError("No image bytes provided")
            image = imdecode_color(image_bytes)
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

    @staticmethod
    def auto_process_image(image_bytes: bytes, max_width: int = 1920, max_height: int = 1080, quality: int = 85) -> np.ndarray:
        """bytes -> BGR image no larger than (max_width, max_height): decode, then Pillow's LANCZOS thumbnail
        (image_processor.py:183-229).  The reference swaps to RGB for Pillow and back; the filter treats every channel
        alike, so the BGR image is resampled directly."""
        try:
            image = imdecode_color(image_bytes)
            if image is None:
                raise ValueError("Failed to decode image")
            return pil_thumbnail_lanczos(image, max_width, max_height)
        except Exception as e:
            logger.error(f"Image processing failed: {str(e)}")
            raise ValueError(f"Image processing error: {str(e)}")
