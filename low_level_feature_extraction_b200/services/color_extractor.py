"""Drop-in for the reference's ColorExtractor (the /extract-colors service).

Mirrors app/services/analyze/color_extractor.py:37-300: same static methods,
signatures, return model and never-raise behaviour of `extract_colors`.

What runs where
  GPU (libllfe.so): BGR->RGB, noise add + clip, np.unique(axis=0) as a 2^24-bit
      bitmap + ordered compaction, and the whole cv2.kmeans (kmeans++ seeding with
      cv::RNG, 10 attempts, Lloyd, empty-cluster repair, best compactness).
  host: drawing the noise tensor from NumPy's global generator (so that
      `np.random.seed(s)` reproduces the reference bit for bit), and the palette
      tail (:231-284: bincount / argsort / hex / filtering), kept as the literal
      NumPy calls because argsort's tie order is platform-defined.

Seeding: the reference draws from two process-global generators.  NumPy's is
used directly.  OpenCV's thread-local cv::RNG cannot be read from Python, so this
module keeps its own copy of that state: `ColorExtractor.set_rng_seed(s)` is the
equivalent of `cv2.setRNGSeed(s)`, and every k-means call advances the state by
exactly the number of draws cv2 would have made.
"""
from __future__ import annotations

import traceback
from typing import Union

import numpy as np
from PIL import Image

from . import _runtime
from .models import ColorFeatures

_CV_RNG_COEFF = 4164903690
_MASK64 = 0xFFFFFFFFFFFFFFFF


class ColorPalette(ColorFeatures):
    """Legacy alias kept for backward compatibility (color_extractor.py:17-35)."""


class ColorExtractor:
    # cv2.kmeans arguments of color_extractor.py:192-196
    KMEANS_ATTEMPTS = 10
    KMEANS_MAX_ITER = 200
    KMEANS_EPS = 0.2
    # "numpy": host noise from np.random (bit-compatible with the reference under np.random.seed);
    # "device": counter-based noise of the same distribution generated in the kernel (throughput mode)
    noise_mode = "numpy"
    _rng_state = 0xFFFFFFFF      # cv::RNG default state (what cv2.setRNGSeed(0) gives)
    _device_seed = 0
    last_status = 0              # LLFE_KMEANS_* bits of the last k-means call (1 = long float32 sums reproduced)

    # ---- seeding ----------------------------------------------------------------
    @classmethod
    def set_rng_seed(cls, seed: int) -> None:
        """Equivalent of cv2.setRNGSeed(seed) for the k-means++ draws."""
        cls._rng_state = (int(seed) & _MASK64) or 0xFFFFFFFF

    @classmethod
    def _advance_rng(cls, draws: int) -> None:
        s = cls._rng_state
        for _ in range(draws):
            s = ((s & 0xFFFFFFFF) * _CV_RNG_COEFF + (s >> 32)) & _MASK64
        cls._rng_state = s

    # ---- pure helpers (color_extractor.py:38-71) ---------------------------------------
    @staticmethod
    def rgb_to_hex(rgb: tuple) -> str:
        return "#{:02x}{:02x}{:02x}".format(rgb[0], rgb[1], rgb[2])

    @staticmethod
    def hex_to_rgb(hex_color: str) -> tuple:
        hex_color = hex_color.lstrip("#")
        return tuple(int(hex_color[i:i + 2], 16) for i in (0, 2, 4))

    @staticmethod
    def get_contrast_ratio(color1: str, color2: str) -> float:
        def get_luminance(c: str) -> float:
            r, g, b = (int(c[i:i + 2], 16) / 255.0 for i in (1, 3, 5) if len(c) >= 6)
            r = r / 12.92 if r <= 0.03928 else ((r + 0.055) / 1.055) ** 2.4
            g = g / 12.92 if g <= 0.03928 else ((g + 0.055) / 1.055) ** 2.4
            b = b / 12.92 if b <= 0.03928 else ((b + 0.055) / 1.055) ** 2.4
            return 0.2126 * r + 0.7152 * g + 0.0722 * b

        l1, l2 = get_luminance(color1), get_luminance(color2)
        lighter, darker = (l1, l2) if l1 > l2 else (l2, l1)
        return (lighter + 0.05) / (darker + 0.05)

    @staticmethod
    def is_light_color(rgb: tuple) -> bool:
        r, g, b = [x / 255.0 for x in rgb]
        return 0.2126 * r + 0.7152 * g + 0.0722 * b > 0.6

    # ---- input normalisation (color_extractor.py:74-171) --------------------------------
    @staticmethod
    def _process_image(image: Union[np.ndarray, Image.Image, None]) -> np.ndarray:
        """Any supported input -> (H, W, 3) uint8 RGB; unsupported / empty -> 100x100 black."""
        default = np.zeros((100, 100, 3), dtype=np.uint8)
        if image is None:
            return default
        if isinstance(image, Image.Image):
            try:
                arr = np.array(image)
                if arr.size == 0:
                    return default
                if image.mode == "RGBA":
                    background = Image.new("RGB", image.size, (255, 255, 255))
                    background.paste(image, mask=image.split()[3])
                    arr = np.array(background)
                elif image.mode != "RGB":
                    arr = np.array(image.convert("RGB"))
                return arr.astype(np.uint8)
            except Exception as e:  # pragma: no cover - mirrors the reference's catch-all
                print(f"Error processing PIL Image: {e}")
                return default
        if isinstance(image, np.ndarray):
            try:
                if image.size == 0:
                    return default
                img = image.copy()
                if img.ndim == 0:
                    return default
                if img.ndim == 1:
                    side = int(np.sqrt(len(img) / 3))
                    if side * side * 3 != len(img):
                        return default
                    img = img.reshape((side, side, 3))
                if img.ndim == 2:
                    img = np.stack([img, img, img], axis=-1)                 # GRAY2RGB
                elif img.ndim == 3:
                    if img.shape[0] <= 4:                                    # the reference's (C,H,W) guess
                        img = np.transpose(img, (1, 2, 0))
                    if img.shape[2] == 1:
                        img = np.repeat(img, 3, axis=2)                      # GRAY2RGB
                    elif img.shape[2] == 3:
                        img = img[..., ::-1]                                 # BGR2RGB
                    elif img.shape[2] == 4:
                        img = img[..., 2::-1]                                # BGRA2RGB (alpha dropped)
                    else:
                        img = img[..., :3]
                if img.dtype != np.uint8:
                    if np.issubdtype(img.dtype, np.floating):
                        img = (img * 255).clip(0, 255).astype(np.uint8)
                    else:
                        img = img.astype(np.uint8)
                return np.ascontiguousarray(img)
            except Exception as e:  # pragma: no cover
                print(f"Error processing numpy array: {e}")
                return default
        return default

    # ---- GPU core ------------------------------------------------------------------------
    @classmethod
    def _dominant_from_rgb(cls, rgb: np.ndarray, noise: np.ndarray | None, n_colors: int):
        """(H,W,3) RGB u8 [+ int8 noise] -> (centers (K,3) u8, labels (U,) int) via libllfe.so."""
        return cls._dominant_from_bgr(np.ascontiguousarray(rgb[..., ::-1]), noise, n_colors)

    @classmethod
    def _dominant_from_bgr(cls, bgr: np.ndarray, noise: np.ndarray | None, n_colors: int):
        """(H,W,3) BGR u8, C-contiguous [+ int8 noise in RGB order] -> (centers (K,3) u8 RGB, labels (U,) int)."""
        h, w = bgr.shape[:2]
        k_req = int(n_colors)
        centers = np.zeros((max(k_req, 1), 3), np.float32)
        labels = np.empty(min(h * w, 1 << 24), np.int32)
        n_unique = np.zeros(1, np.int32)
        k_used = np.zeros(1, np.int32)
        comp = np.zeros(1, np.float64)
        status = np.zeros(1, np.int32)
        # the unique list itself is what the reference returns when fewer than two clusters are asked for / possible
        keys = np.empty(labels.shape, np.uint32) if k_req <= 1 else None
        nz = None if noise is None else np.ascontiguousarray(noise.reshape(h, w, 3))
        with _runtime.lock():
            cls._device_seed += 1
            _runtime.context().call("llfe_dominant_colors_host", bgr, h, w, nz, cls._device_seed, max(k_req, 1),
                                    cls.KMEANS_ATTEMPTS, cls.KMEANS_MAX_ITER, cls.KMEANS_EPS, cls._rng_state,
                                    centers, labels, n_unique, k_used, comp, keys, status)
        u, k = int(n_unique[0]), int(k_used[0])
        cls.last_status = int(status[0])
        actual = min(k_req, u)
        if actual < k_req:
            print(f"Warning: Only {u} unique colors found, reducing number of clusters from {k_req} to {actual}")
        if actual <= 1:
            # color_extractor.py:185-186: `return unique_colors, np.array([0] * len(unique_colors))` -- ALL unique
            # colours as "centres" (also for n_colors <= 1 on a many-coloured image), labels all zero
            if keys is None:      # n_colors >= 2 but the image has a single colour: the centre is that colour
                return centers[:u].astype(np.uint8), np.zeros(u, np.int64)
            kk = keys[:u]
            uniq = np.stack([(kk >> 16) & 255, (kk >> 8) & 255, kk & 255], axis=1).astype(np.uint8)
            return uniq, np.zeros(u, np.int64)
        if k > 1:   # cv2 made 1 + 6 (K-1) draws per attempt
            cls._advance_rng(cls.KMEANS_ATTEMPTS * (1 + 6 * (k - 1)))
        # color_extractor.py:197 truncates float centres with astype(uint8)
        return centers[:k].astype(np.uint8), labels[:u].astype(np.int64)

    @staticmethod
    def _get_dominant_colors(pixels: np.ndarray, n_colors: int) -> tuple:
        """(N,3) uint8 RGB pixel list -> (centers, labels)   (color_extractor.py:173-201)."""
        px = np.ascontiguousarray(pixels, dtype=np.uint8).reshape(-1, 3)
        if len(px) == 0:
            return np.zeros((0, 3), np.uint8), np.array([], dtype=np.int64)
        zero = np.zeros(px.shape, np.int8)
        return ColorExtractor._dominant_from_rgb(px.reshape(1, -1, 3), zero, n_colors)

    @staticmethod
    def _palette_from_clusters(centers: np.ndarray, counts: np.ndarray | None) -> ColorFeatures:
        """color_extractor.py:231-284 on host: order the (K,3) u8 RGB centres by the number of distinct colours
        in their cluster (the literal NumPy argsort), drop pure white / black, pick primary, three
        accents and the contrasting background."""
        if len(centers) > 1:
            sorted_indices = np.argsort(-counts)
            centers = centers[sorted_indices]
            counts = counts[sorted_indices]

        hex_colors = [ColorExtractor.rgb_to_hex(tuple(int(v) for v in color)) for color in centers]
        hex_colors = [c for c in hex_colors if c.lower() not in ["#ffffff", "#000000"]]
        meta = {"success": True, "timestamp": 0.0, "processing_time": 0.0}
        if not hex_colors:
            bg_color = "#000000" if ColorExtractor.is_light_color((255, 255, 255)) else "#FFFFFF"
            return ColorFeatures(primary=bg_color, background=bg_color, accent=[bg_color] * 3, metadata=meta)
        primary = hex_colors[0]
        accent_colors = [c for c in hex_colors if c != primary][:3]
        while len(accent_colors) < 3:
            accent_colors.append(accent_colors[-1] if accent_colors else primary)
        bg_color = "#FFFFFF" if not ColorExtractor.is_light_color(ColorExtractor.hex_to_rgb(primary)) else "#000000"
        return ColorFeatures(primary=primary, background=bg_color, accent=accent_colors[:3], metadata=meta)

    @staticmethod
    def _palettes_from_batch(centers: np.ndarray, k_used: np.ndarray, sizes: np.ndarray) -> list:
        """`_palette_from_clusters` for a whole batch: centers (n, K, 3) float32 as cv2.kmeans returns them, k_used
        (n,), sizes (n, K) = np.bincount(labels) per image.  The ordering is the same literal NumPy call
        (`np.argsort(-counts)`, applied row by row by `axis=1`: NumPy runs its 1-D argsort on every row, so ties fall
        exactly as in the per-image call); everything else is the reference's arithmetic on integers instead of
        on strings.  Equality with the per-image function is tested on the host (tests/test_services_host.py)."""
        n, kk = sizes.shape
        c8 = centers.astype(np.uint8)                                   # color_extractor.py:197 truncation
        out: list = [None] * n
        full = np.flatnonzero(k_used == kk) if kk > 1 else np.empty(0, np.int64)
        if len(full):
            order = np.argsort(-sizes[full].astype(np.int64), axis=1)  # color_extractor.py:234, one row per image
            cs = np.take_along_axis(c8[full], order[:, :, None], axis=1).astype(np.uint32)
            codes = ((cs[:, :, 0] << 16) | (cs[:, :, 1] << 8) | cs[:, :, 2]).tolist()
            meta = {"success": True, "timestamp": 0.0, "processing_time": 0.0}
            for row, i in zip(codes, full.tolist()):
                cols = [v for v in row if v != 0xFFFFFF and v != 0]       # :242 drops pure white / black
                if not cols:
                    bg = "#000000" if ColorExtractor.is_light_color((255, 255, 255)) else "#FFFFFF"
                    out[i] = ColorFeatures(primary=bg, background=bg, accent=[bg] * 3, metadata=dict(meta))
                    continue
                p = cols[0]
                acc = [v for v in cols if v != p][:3]
                while len(acc) < 3:
                    acc.append(acc[-1] if acc else p)
                light = ColorExtractor.is_light_color((p >> 16, (p >> 8) & 255, p & 255))
                out[i] = ColorFeatures(primary="#%06x" % p, background="#000000" if light else "#FFFFFF",
                                       accent=["#%06x" % v for v in acc], metadata=dict(meta))
        for i in range(n):
            if out[i] is None:      # fewer colours than clusters (or K = 1): the per-image path
                k = int(k_used[i])
                out[i] = ColorExtractor._palette_from_clusters(c8[i, :k], sizes[i, :k].astype(np.int64) if k > 1 else None)
        return out

    # ---- the service call -------------------------------------------------------------------
    @staticmethod
    def extract_colors(image: Union[np.ndarray, Image.Image], n_colors: int = 5) -> ColorFeatures:
        try:
            plain_bgr = (isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3 and
                         image.shape[2] == 3 and image.shape[0] > 4 and image.size > 0)
            if plain_bgr:
                # the common case (a decoded BGR frame): _process_image would only swap the channels, and the
                # kernels read BGR directly -- skip the two host-side channel reversals (6 MB copies at 1080p)
                bgr = np.ascontiguousarray(image)
                pixels = bgr.reshape(-1, 3)
            else:
                bgr = None
                img_array = ColorExtractor._process_image(image)
                pixels = img_array.reshape(-1, 3)
            if len(pixels) > 0:
                if ColorExtractor.noise_mode == "numpy":
                    noise = np.random.normal(0, 0.5, pixels.shape).astype(np.int8)   # color_extractor.py:224
                else:
                    noise = None
                if bgr is not None:
                    centers, labels = ColorExtractor._dominant_from_bgr(bgr, noise, n_colors)
                else:
                    centers, labels = ColorExtractor._dominant_from_rgb(img_array, noise, n_colors)
            else:
                centers, labels = np.zeros((0, 3), np.uint8), np.array([], dtype=np.int64)

            counts = np.bincount(labels, minlength=len(centers)) if len(centers) > 1 else None
            return ColorExtractor._palette_from_clusters(centers, counts)
        except Exception as e:
            print(f"Error in extract_colors: {str(e)}\n{traceback.format_exc()}")
            return ColorFeatures(primary="#000000", background="#FFFFFF", accent=["#666666", "#999999", "#CCCCCC"],
                                 metadata={"success": False, "error": str(e), "timestamp": 0.0, "processing_time": 0.0})
