"""Device-resident batched ops: torch CUDA tensors in, torch CUDA tensors out.

torch is plumbing here (device memory, streams, torch.distributed); every op
is a call through the C ABI of libllfe.so on torch's current stream.  Inputs
are uint8 (n, h, w, 3) BGR batches (or (h, w, 3) single images).
"""
from __future__ import annotations

import threading

import torch

from ._native import Context

_engines: dict[int, "Engine"] = {}
_lock = threading.Lock()


def engine(device: int | torch.device | None = None) -> "Engine":
    """The per-device Engine singleton (one llfe context per process and device)."""
    if device is None:
        idx = torch.cuda.current_device()
    elif isinstance(device, torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
    else:
        idx = int(device)
    with _lock:
        if idx not in _engines:
            _engines[idx] = Engine(idx)
        return _engines[idx]


def _batch(x: torch.Tensor, channels: int | None):
    """-> (contiguous batched tensor, was_single)."""
    if x.dtype != torch.uint8 or not x.is_cuda:
        raise TypeError("expected a CUDA uint8 tensor")
    if channels is None:  # (n,h,w) or (h,w)
        single = x.dim() == 2
        if x.dim() not in (2, 3):
            raise ValueError(f"expected (n,h,w) or (h,w), got {tuple(x.shape)}")
    else:
        single = x.dim() == 3
        if x.dim() not in (3, 4) or x.shape[-1] != channels:
            raise ValueError(f"expected (n,h,w,{channels}) or (h,w,{channels}), got {tuple(x.shape)}")
    if single:
        x = x.unsqueeze(0)
    return x.contiguous(), single


class Engine:
    def __init__(self, device_index: int):
        if not torch.cuda.is_available():
            raise RuntimeError("low_level_feature_extraction_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", device_index)
        self.ctx = Context(device_index)

    # -- plumbing -------------------------------------------------------------
    def _bind(self):
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def _empty(self, shape, dtype=torch.uint8):
        return torch.empty(shape, dtype=dtype, device=self.device)

    @property
    def launches(self) -> int:
        return self.ctx.launches

    # -- pointwise --------------------------------------------------------------
    def bgr2gray(self, bgr: torch.Tensor) -> torch.Tensor:
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_bgr2gray", x, n, h, w, out)
        return out[0] if single else out

    def bgr2rgb(self, bgr: torch.Tensor) -> torch.Tensor:
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        out = self._empty((n, h, w, 3))
        self._bind()
        self.ctx.call("llfe_bgr2rgb", x, n, h, w, out)
        return out[0] if single else out

    def convert_scale_abs(self, src: torch.Tensor, alpha: float, alpha2: float | None = None) -> torch.Tensor:
        if src.dtype != torch.uint8 or not src.is_cuda:
            raise TypeError("expected a CUDA uint8 tensor")
        x = src.contiguous()
        out = torch.empty_like(x)
        self._bind()
        self.ctx.call("llfe_convert_scale_abs", x, x.numel(), float(alpha), float(alpha2 if alpha2 is not None else 1.0),
                      1 if alpha2 is None else 0, out)
        return out

    # -- blur --------------------------------------------------------------------
    def gaussian_blur5(self, src: torch.Tensor) -> torch.Tensor:
        if src.dim() >= 3 and src.shape[-1] == 3:
            x, single = _batch(src, 3)
            n, h, w, c = x.shape
        else:
            x, single = _batch(src, None)
            n, h, w = x.shape
            c = 1
        out = torch.empty_like(x)
        self._bind()
        self.ctx.call("llfe_gaussian_blur5", x, n, h, w, c, out)
        return out[0] if single else out

    def gray_blur5(self, bgr: torch.Tensor) -> torch.Tensor:
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_gray_blur5", x, n, h, w, out)
        return out[0] if single else out

    # -- edges -------------------------------------------------------------------
    def canny(self, gray: torch.Tensor, low: int = 50, high: int = 150) -> torch.Tensor:
        x, single = _batch(gray, None)
        n, h, w = x.shape
        out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_canny", x, n, h, w, int(low), int(high), out)
        return out[0] if single else out

    def dilate3(self, src: torch.Tensor) -> torch.Tensor:
        x, single = _batch(src, None)
        n, h, w = x.shape
        out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_dilate3", x, n, h, w, out)
        return out[0] if single else out

    def shape_mask(self, bgr: torch.Tensor, low: int = 50, high: int = 150, out: torch.Tensor | None = None):
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        if out is None:
            out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_shape_mask", x, n, h, w, int(low), int(high), out)
        return out[0] if single and out.dim() == 3 else out

    # -- thresholds ---------------------------------------------------------------
    def adaptive_threshold(self, gray: torch.Tensor, c: int = 2, with_sums: bool = False):
        x, single = _batch(gray, None)
        n, h, w = x.shape
        out = self._empty((n, h, w))
        sums = self._empty((n, 2), torch.int64) if with_sums else None
        self._bind()
        self.ctx.call("llfe_adaptive_threshold", x, n, h, w, int(c), out, sums)
        if with_sums:
            return (out[0], sums[0]) if single else (out, sums)
        return out[0] if single else out

    def shadow_mask(self, bgr: torch.Tensor, want_blurred: bool = False):
        """-> (mask, sum_count (n,2) int64[, blurred])."""
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        mask = self._empty((n, h, w))
        blurred = self._empty((n, h, w)) if want_blurred else None
        sums = self._empty((n, 2), torch.int64)
        self._bind()
        self.ctx.call("llfe_shadow_mask", x, n, h, w, mask, blurred, sums)
        res = (mask, sums, blurred) if want_blurred else (mask, sums)
        return tuple(t[0] for t in res) if single else res

    def font_mask(self, bgr: torch.Tensor) -> torch.Tensor:
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_font_mask", x, n, h, w, out)
        return out[0] if single else out

    def otsu(self, gray: torch.Tensor, invert_if_light: bool = False):
        """-> (mask, thresholds int32 (n,))."""
        x, single = _batch(gray, None)
        n, h, w = x.shape
        out = self._empty((n, h, w))
        thr = self._empty((n,), torch.int32)
        self._bind()
        self.ctx.call("llfe_otsu", x, n, h, w, 1 if invert_if_light else 0, out, thr)
        return (out[0], thr[0]) if single else (out, thr)

    def text_mask(self, bgr: torch.Tensor):
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        out = self._empty((n, h, w))
        thr = self._empty((n,), torch.int32)
        self._bind()
        self.ctx.call("llfe_text_mask", x, n, h, w, out, thr)
        return (out[0], thr[0]) if single else (out, thr)

    # -- resize -------------------------------------------------------------------
    def resize_area(self, src: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
        if src.dim() >= 3 and src.shape[-1] == 3:
            x, single = _batch(src, 3)
            n, sh, sw, c = x.shape
            out = self._empty((n, dh, dw, 3))
        else:
            x, single = _batch(src, None)
            n, sh, sw = x.shape
            c = 1
            out = self._empty((n, dh, dw))
        self._bind()
        self.ctx.call("llfe_resize_area", x, n, sh, sw, c, out, int(dh), int(dw))
        return out[0] if single else out
