"""Drop-in for the hot-path part of the reference's ImageTransformer
(app/services/__pycache__/image_transformer.cpython-312.pyc, src L9-176):
ndarray branches of `resize` / `generate_thumbnail` (INTER_AREA), `apply_filter`
('gaussian_blur') and `adjust_brightness_contrast` (two convertScaleAbs fused
into one LUT pass).  PIL branches and the box / median filters are outside the
path and keep the reference's own library calls."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import cv2
import numpy as np
from PIL import Image, ImageEnhance, ImageFilter

from . import _runtime
from .image_processor import resize_area


class ImageTransformer:
    @staticmethod
    def resize(image: Union[Image.Image, np.ndarray], width: Optional[int] = None, height: Optional[int] = None,
               maintain_aspect_ratio: bool = True) -> Union[Image.Image, np.ndarray]:
        """src L10-59."""
        if isinstance(image, Image.Image):
            if maintain_aspect_ratio:
                image.thumbnail((width or height, height or width))
                return image
            return image.resize((width, height))
        if isinstance(image, np.ndarray):
            orig_height, orig_width = image.shape[:2]
            if maintain_aspect_ratio:
                scale = min((width or orig_width) / orig_width, (height or orig_height) / orig_height)
                new_width = int(orig_width * scale)
                new_height = int(orig_height * scale)
            else:
                new_width = width or orig_width
                new_height = height or orig_height
            return resize_area(image, new_width, new_height)
        raise TypeError("Unsupported image type")

    @staticmethod
    def apply_filter(image: Union[Image.Image, np.ndarray], filter_type: str = "blur") -> Union[Image.Image, np.ndarray]:
        """src L61-109."""
        if isinstance(image, Image.Image):
            filters = {"blur": ImageFilter.BLUR, "contour": ImageFilter.CONTOUR, "detail": ImageFilter.DETAIL,
                       "edge_enhance": ImageFilter.EDGE_ENHANCE, "emboss": ImageFilter.EMBOSS,
                       "sharpen": ImageFilter.SHARPEN}
            if filter_type not in filters:
                raise ValueError(f"Unsupported filter type: {filter_type}")
            return image.filter(filters[filter_type])
        if isinstance(image, np.ndarray):
            if filter_type == "gaussian_blur":
                src = np.ascontiguousarray(image)
                if src.dtype != np.uint8 or src.ndim not in (2, 3) or (src.ndim == 3 and src.shape[2] != 3):
                    raise TypeError("Unsupported image type")
                h, w = src.shape[:2]
                dst = np.empty_like(src)
                with _runtime.lock():
                    _runtime.context().call("llfe_gaussian_blur5_host", src, h, w, 1 if src.ndim == 2 else 3, dst)
                return dst
            if filter_type == "blur":            # 3x3 box: not on the hot path
                return cv2.blur(image, (3, 3))
            if filter_type == "median_blur":     # not on the hot path
                return cv2.medianBlur(image, 3)
            raise ValueError(f"Unsupported filter type: {filter_type}")
        raise TypeError("Unsupported image type")

    @staticmethod
    def adjust_brightness_contrast(image: Union[Image.Image, np.ndarray], brightness: float = 1.0,
                                   contrast: float = 1.0) -> Union[Image.Image, np.ndarray]:
        """src L111-146: convertScaleAbs(alpha=brightness) then convertScaleAbs(alpha=contrast)."""
        if isinstance(image, Image.Image):
            image = ImageEnhance.Brightness(image).enhance(brightness)
            return ImageEnhance.Contrast(image).enhance(contrast)
        if isinstance(image, np.ndarray):
            src = np.ascontiguousarray(image)
            if src.dtype != np.uint8:
                raise TypeError("Unsupported image type")
            dst = np.empty_like(src)
            with _runtime.lock():
                _runtime.context().call("llfe_convert_scale_abs_host", src, src.size, float(brightness),
                                        float(contrast), 0, dst)
            return dst
        raise TypeError("Unsupported image type")

    @staticmethod
    def generate_thumbnail(image: Union[Image.Image, np.ndarray], size: Tuple[int, int] = (128, 128)):
        """src L148-176."""
        if isinstance(image, Image.Image):
            image.thumbnail(size)
            return image
        if isinstance(image, np.ndarray):
            return resize_area(image, size[0], size[1])
        raise TypeError("Unsupported image type")
