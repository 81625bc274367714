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

    def hysteresis(self, weak: torch.Tensor, strong: torch.Tensor, dilate: bool = False) -> torch.Tensor:
        """Canny's hysteresis stage on u8 maps (non-zero = set): weak pixels connected to a strong pixel."""
        x, single = _batch(weak, None)
        s, _ = _batch(strong, None)
        if s.shape != x.shape:
            raise ValueError("weak and strong must have the same shape")
        n, h, w = x.shape
        out = self._empty((n, h, w))
        self._bind()
        self.ctx.call("llfe_hysteresis", x, s, n, h, w, 1 if dilate else 0, out)
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
    def contours_external(self, mask: torch.Tensor, min_area2: int = 200, max_contours: int = 4096,
                          max_points: int = 1 << 16):
        """cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) per image on the device.
        -> headers (n, max_contours, 10) int32 [contours.HEADER records], points (n, max_points, 2) int32,
        counts (n, 4) int32; see include/llfe.h for the overflow protocol."""
        x, single = _batch(mask, None)
        n, h, w = x.shape
        hdr = self._empty((n, max_contours, 10), torch.int32)
        pts = self._empty((n, max(max_points, 1), 2), torch.int32)
        cnt = self._empty((n, 4), torch.int32)
        self._bind()
        self.ctx.call("llfe_contours_external", x, n, h, w, int(min_area2), hdr, max_contours, pts, max_points, cnt)
        return (hdr[0], pts[0], cnt[0]) if single else (hdr, pts, cnt)

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

    def resize_linear(self, src: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
        """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) on uint8 (1 or 3 channels)."""
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
        self.ctx.call("llfe_resize_linear", x, n, sh, sw, c, out, int(dh), int(dw))
        return out[0] if single else out

    def resize_lanczos4(self, src: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
        """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4) on uint8 (1 or 3 channels)."""
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
        self.ctx.call("llfe_resize_lanczos4", x, n, sh, sw, c, out, int(dh), int(dw))
        return out[0] if single else out

    # -- decode / Pillow thumbnail --------------------------------------
    def png_reconstruct(self, streams: torch.Tensor, h: int, w: int, color_type: int, bit_depth: int,
                        palettes: torch.Tensor | None = None):
        """Inflated IDAT streams (n, h * (1 + rowbytes)) uint8, reconstructed IN PLACE -> (BGR images (n,h,w,3),
        status int32 (n,): non-zero = invalid filter byte).  palettes: (n, 256, 3) RGB for colour type 3."""
        rb = int(self.ctx.lib.llfe_png_rowbytes(int(w), int(color_type), int(bit_depth)))
        if rb <= 0:
            raise ValueError("invalid PNG colour type / bit depth")
        if streams.dtype != torch.uint8 or not streams.is_cuda or streams.dim() != 2 or streams.shape[1] != h * (rb + 1) \
                or not streams.is_contiguous():
            raise ValueError(f"expected a contiguous CUDA uint8 tensor (n, {h * (rb + 1)})")
        n = streams.shape[0]
        if color_type == 3:
            if palettes is None or palettes.dtype != torch.uint8 or tuple(palettes.shape) != (n, 256, 3):
                raise ValueError("colour type 3 needs palettes (n, 256, 3) uint8")
            palettes = palettes.contiguous()
        out = self._empty((n, h, w, 3))
        status = self._empty((n,), torch.int32)
        self._bind()
        self.ctx.call("llfe_png_reconstruct", streams, n, int(h), int(w), int(color_type), int(bit_depth), palettes, out, status)
        return out, status

    def pil_reduce(self, src: torch.Tensor, fx: int, fy: int, box=None) -> torch.Tensor:
        """Pillow's Image.reduce((fx, fy), box) on uint8 (1 or 3 channels)."""
        import numpy as np

        if src.dim() >= 3 and src.shape[-1] == 3:
            x, single = _batch(src, 3)
            n, sh, sw, c = x.shape
        else:
            x, single = _batch(src, None)
            n, sh, sw = x.shape
            c = 1
        b = np.asarray(box if box is not None else (0, 0, sw, sh), np.int32)
        dh, dw = -(-int(b[3] - b[1]) // fy), -(-int(b[2] - b[0]) // fx)
        out = self._empty((n, dh, dw, 3) if c == 3 else (n, dh, dw))
        self._bind()
        self.ctx.call("llfe_pil_reduce", x, n, sh, sw, c, b, int(fx), int(fy), out)
        return out[0] if single else out

    def pil_resample_lanczos(self, src: torch.Tensor, dh: int, dw: int, box=None) -> torch.Tensor:
        """Pillow's im.resize((dw, dh), LANCZOS, box, reducing_gap=None) on uint8 (1 or 3 channels)."""
        import numpy as np

        if src.dim() >= 3 and src.shape[-1] == 3:
            x, single = _batch(src, 3)
            n, sh, sw, c = x.shape
        else:
            x, single = _batch(src, None)
            n, sh, sw = x.shape
            c = 1
        b = np.asarray(box if box is not None else (0, 0, sw, sh), np.float32)
        out = self._empty((n, dh, dw, 3) if c == 3 else (n, dh, dw))
        self._bind()
        self.ctx.call("llfe_pil_resample_lanczos", x, n, sh, sw, c, b, out, int(dh), int(dw))
        return out[0] if single else out

    # -- palette ------------------------------------------------------------------
    def unique_colors(self, bgr: torch.Tensor, noise: torch.Tensor | None = None, seed: int = 0,
                      max_unique: int = 1 << 16, with_counts: bool = False, first_image: int = 0):
        """np.unique(noised RGB pixels, axis=0) per image.

        noise: int8 (n,h,w,3) tensor in RGB order (the reference's
        `np.random.normal(0, 0.5, pixels.shape).astype(np.int8)`), or None to
        generate noise of the same distribution on the device from `seed` (a pure function of
        seed, first_image + i and the pixel position).
        -> (keys uint32-as-int32 (n,max_unique) = R<<16|G<<8|B ascending,
            count int32 (n,)[, pixel counts int32 (n,max_unique)])."""
        x, single = _batch(bgr, 3)
        n, h, w, _ = x.shape
        if noise is not None:
            if noise.dtype != torch.int8 or noise.numel() != x.numel():
                raise ValueError("noise must be int8 with the shape of the image batch")
            noise = noise.contiguous()
        keys = self._empty((n, max_unique), torch.int32)
        count = self._empty((n,), torch.int32)
        hist = self._empty((n, max_unique), torch.int32) if with_counts else None
        self._bind()
        self.ctx.call("llfe_unique_colors", x, n, h, w, noise, int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_image), keys,
                      hist, count, int(max_unique))
        res = (keys, count, hist) if with_counts else (keys, count)
        return tuple(t[0] for t in res) if single else res

    def kmeans_unique(self, keys: torch.Tensor, count: torch.Tensor, k: int, rng_state, attempts: int = 10,
                      max_iter: int = 200, eps: float = 0.2):
        """cv2.kmeans(float32(unique), k, None, (EPS+MAX_ITER, max_iter, eps), attempts, KMEANS_PP_CENTERS)
        per image.  rng_state: int or sequence of ints (cv::RNG state; cv2.setRNGSeed(s) => s).
        -> (centers float32 (n,k,3) RGB, labels int32 (n,max_unique), compactness float64 (n,), k_used int32 (n,)).
        self.last_cluster_sizes holds np.bincount(labels) per image (int32 (n,k)), self.last_status the
        LLFE_KMEANS_* bits per image (1 = long float32 sums reproduced, 2 = list truncated: k_used = -1)."""
        if keys.dim() == 1:
            keys, count = keys.unsqueeze(0), count.reshape(1)
        n, max_unique = keys.shape
        states = [rng_state] * n if isinstance(rng_state, int) else list(rng_state)
        rs = torch.tensor([int(s) & 0xFFFFFFFFFFFFFFFF for s in states], dtype=torch.uint64).view(torch.int64).to(self.device)
        centers = torch.zeros((n, k, 3), dtype=torch.float32, device=self.device)
        labels = self._empty((n, max_unique), torch.int32)
        comp = self._empty((n,), torch.float64)
        kused = self._empty((n,), torch.int32)
        sizes = self._empty((n, k), torch.int32)
        status = self._empty((n,), torch.int32)
        self._bind()
        self.ctx.call("llfe_kmeans_unique", keys.contiguous(), count.contiguous(), n, max_unique, int(k), int(attempts),
                      int(max_iter), float(eps), rs, centers, labels, comp, kused, sizes, status)
        self.last_cluster_sizes = sizes
        self.last_status = status
        return centers, labels, comp, kused

    def palette_large(self, bgr_image: torch.Tensor, n_unique: int, k: int, rng_state: int, seed: int = 0,
                      first_image: int = 0, noise: torch.Tensor | None = None, attempts: int = 10, max_iter: int = 200,
                      eps: float = 0.2):
        """The palette of ONE image whose unique-colour list has n_unique entries (any length up to 2^24): the
        list is sized from n_unique instead of a batch-wide capacity.  Same noise as the batch call that counted the
        colours when (seed, first_image) are that call's seed and the image's index in it.
        -> (centers (k,3) f32, k_used (1,) i32, cluster_sizes (k,) i32, status (1,) i32) device tensors."""
        cap = max(1, int(n_unique))
        keys, count = self.unique_colors(bgr_image, noise, seed=seed, max_unique=cap, first_image=first_image)
        centers, _, _, kused = self.kmeans_unique(keys, count, k, rng_state, attempts, max_iter, eps)
        return centers[0], kused, self.last_cluster_sizes[0], self.last_status

    def kmeans_lloyd(self, keys: torch.Tensor, count: torch.Tensor, init_centers: torch.Tensor,
                     weights: torch.Tensor | None = None, exact_sums: bool = False, max_iter: int = 200,
                     eps: float = 0.2):
        """Lloyd from given centres over (optionally weighted) colour lists.
        -> (centers (n,k,3) f32, labels int32 (n,max_unique), iters int32 (n,), sums_counts int64 (n,k,4))."""
        if keys.dim() == 1:
            keys, count = keys.unsqueeze(0), count.reshape(1)
            init_centers = init_centers.unsqueeze(0)
            if weights is not None:
                weights = weights.unsqueeze(0)
        n, max_unique = keys.shape
        k = init_centers.shape[1]
        init = init_centers.to(device=self.device, dtype=torch.float32).contiguous()
        centers = torch.zeros((n, k, 3), dtype=torch.float32, device=self.device)
        labels = self._empty((n, max_unique), torch.int32)
        iters = self._empty((n,), torch.int32)
        sums = torch.zeros((n, k, 4), dtype=torch.int64, device=self.device)
        self._bind()
        self.ctx.call("llfe_kmeans_lloyd", keys.contiguous(), weights.contiguous() if weights is not None else None,
                      count.contiguous(), n, max_unique, k, int(max_iter), float(eps), 1 if exact_sums else 0, init,
                      centers, labels, iters, sums)
        return centers, labels, iters, sums

    # -- per-pixel k-means building blocks (row shard of one image) ---------------
    def kmeans_pixels_step(self, bgr_rows: torch.Tensor, centers: torch.Tensor, sums: torch.Tensor,
                           labels: torch.Tensor | None = None, state: torch.Tensor | None = None):
        """sums (k,4) int64 += exact {sum R, sum G, sum B, count} per cluster over the pixels of bgr_rows
        (a no-op when `state` says converged or frozen)."""
        x = bgr_rows.contiguous()
        npix = x.numel() // 3
        k = centers.shape[0]
        self._bind()
        self.ctx.call("llfe_kmeans_pixels_step", x, npix, k, centers, sums, labels, state)

    def kmeans_pixels_zero(self, sums: torch.Tensor, state: torch.Tensor | None = None):
        """Zero the (k,4) accumulator unless `state` says converged or frozen."""
        self._bind()
        self.ctx.call("llfe_kmeans_pixels_zero", sums.shape[0], sums, state)

    def kmeans_update(self, sums: torch.Tensor, centers: torch.Tensor, state: torch.Tensor, shift: torch.Tensor,
                      max_iter: int = 200, eps: float = 0.2, consumed: torch.Tensor | None = None,
                      zero_sums: bool = False):
        """Centres from the (all-reduced) sums + convergence bookkeeping in `state`.  `consumed` receives the
        sums that were used; zero_sums clears `sums` afterwards (in-place all-reduce loop, see llfe.h)."""
        self._bind()
        self.ctx.call("llfe_kmeans_update", centers.shape[0], sums, centers, int(max_iter), float(eps), state, shift,
                      consumed, int(bool(zero_sums)))

    # ---- colour-histogram form of the per-pixel k-means (config 5) ----------------------------------
    HIST_BINS = 1 << 24

    def kmeans_update_p2p(self, partial: torch.Tensor, mailboxes: torch.Tensor, rank: int, world: int,
                          centers: torch.Tensor, state: torch.Tensor, shift: torch.Tensor, totals: torch.Tensor,
                          max_iter: int = 200, eps: float = 0.2):
        """All-reduce of the (k,4) partial sums through the peers' mailboxes + centre update in one kernel
        (llfe_kmeans_update_p2p): `partial` is consumed and cleared, `totals` receives the global sums."""
        self._bind()
        self.ctx.call("llfe_kmeans_update_p2p", partial.shape[0], partial, mailboxes, int(rank), int(world), centers,
                      int(max_iter), float(eps), state, shift, totals)

    # -- peer memory (one node, NVLink): raw device allocations that can be exported to the other ranks ----------
    def raw_malloc(self, nbytes: int) -> int:
        import ctypes as C

        ptr = C.c_void_p()
        rc = self.ctx.lib.llfe_malloc(self.ctx.handle, int(nbytes), C.byref(ptr))
        if rc != 0:
            raise RuntimeError(self.ctx.lib.llfe_last_error().decode())
        self._bind()
        self.ctx.call("llfe_memset", ptr.value, 0, int(nbytes))
        return int(ptr.value)

    def raw_memset(self, ptr: int, value: int, nbytes: int):
        self._bind()
        self.ctx.call("llfe_memset", int(ptr), int(value), int(nbytes))

    def pixels_histogram_raw(self, bgr_rows: torch.Tensor, hist_ptr: int):
        """pixels_histogram into a raw (exported) 2^24-bin table."""
        x = bgr_rows.contiguous()
        self._bind()
        self.ctx.call("llfe_pixels_histogram", x, x.numel() // 3, int(hist_ptr))

    def p2p_barrier(self, mailboxes: torch.Tensor, rank: int, world: int):
        self._bind()
        self.ctx.call("llfe_p2p_barrier", mailboxes, int(rank), int(world))

    def histogram_pull_reduce(self, tables: torch.Tensor, rank: int, world: int, share: torch.Tensor):
        self._bind()
        self.ctx.call("llfe_histogram_pull_reduce", tables, int(rank), int(world), share)

    def raw_free(self, ptr: int):
        self.ctx.call("llfe_free", int(ptr))

    def ipc_export(self, ptr: int) -> bytes:
        import ctypes as C

        buf = (C.c_uint8 * 64)()
        self.ctx.call("llfe_ipc_export", int(ptr), C.addressof(buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        import ctypes as C

        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        rc = self.ctx.lib.llfe_ipc_open(self.ctx.handle, C.addressof(buf), C.byref(out))
        if rc != 0:
            raise RuntimeError(self.ctx.lib.llfe_last_error().decode())
        return int(out.value)

    def ipc_close(self, ptr: int):
        self.ctx.call("llfe_ipc_close", int(ptr))

    def pixels_histogram(self, bgr_rows: torch.Tensor, hist: torch.Tensor):
        """hist (2^24,) int32 += pixel count per colour key (R << 16) + (G << 8) + B."""
        x = bgr_rows.contiguous()
        assert hist.numel() == self.HIST_BINS and hist.dtype == torch.int32 and hist.is_contiguous()
        self._bind()
        self.ctx.call("llfe_pixels_histogram", x, x.numel() // 3, hist)

    def histogram_compact(self, hist: torch.Tensor, part: int = 0, parts: int = 1):
        """(keys, counts): the non-empty bins of this part's interleaved blocks, key ascending (int32 tensors)."""
        self._bind()
        n = torch.zeros((1,), dtype=torch.int32, device=hist.device)
        self.ctx.call("llfe_histogram_compact", hist, int(part), int(parts), None, None, 0, n, 0)
        u = int(n.item())
        keys = torch.empty((u,), dtype=torch.int32, device=hist.device)
        counts = torch.empty((u,), dtype=torch.int32, device=hist.device)
        if u:
            self.ctx.call("llfe_histogram_compact", hist, int(part), int(parts), keys, counts, u, n, 0)
        return keys, counts

    def histogram_compact_device(self, hist: torch.Tensor, part: int = 0, parts: int = 1, packed: bool = False,
                                 cap: int | None = None):
        """Like histogram_compact but without the host round trip: (keys, counts) have room for every bin of the
        part (or `cap`), the number of entries stays on the device -> (keys, counts, n (1,) int32).  packed: `hist`
        is this part's share only (what `reduce_scatter` over the block-transposed table returns)."""
        bins = hist.numel() if packed else self.HIST_BINS // int(parts) + 2048
        cap = bins if cap is None else min(int(cap), bins)
        keys = torch.empty((cap,), dtype=torch.int32, device=hist.device)
        counts = torch.empty((cap,), dtype=torch.int32, device=hist.device)
        n = torch.zeros((1,), dtype=torch.int32, device=hist.device)
        self._bind()
        self.ctx.call("llfe_histogram_compact", hist, int(part), int(parts), keys, counts, cap, n, 1 if packed else 0)
        return keys, counts, n

    def kmeans_hist_step(self, keys: torch.Tensor, counts: torch.Tensor, centers: torch.Tensor, sums: torch.Tensor,
                         labels: torch.Tensor | None = None, state: torch.Tensor | None = None,
                         n_dev: torch.Tensor | None = None):
        """sums (k,4) int64 += count * {R, G, B, 1} per cluster over the (key, count) entries (the first n_dev[0] of
        them when the count lives on the device)."""
        self._bind()
        self.ctx.call("llfe_kmeans_hist_step", keys, counts, keys.numel(), centers.shape[0], centers, sums, labels,
                      state, n_dev)

    def kmeans_hist_lloyd(self, keys: torch.Tensor, counts: torch.Tensor, centers: torch.Tensor, partial: torch.Tensor,
                          labels: torch.Tensor | None, state: torch.Tensor, shift: torch.Tensor, totals: torch.Tensor,
                          n_dev: torch.Tensor | None = None, mailboxes: torch.Tensor | None = None, rank: int = 0,
                          world: int = 1, max_iter: int = 200, eps: float = 0.2, iterations: int = 32):
        """Up to `iterations` Lloyd iterations over (key, count) entries in one persistent cooperative kernel, the sums
        exchanged through the peers' mailboxes (llfe_kmeans_hist_lloyd); stops early on converged / frozen."""
        self._bind()
        self.ctx.call("llfe_kmeans_hist_lloyd", keys, counts, keys.numel(), n_dev, centers.shape[0], centers, partial, labels,
                      mailboxes, int(rank), int(world), int(max_iter), float(eps), state, shift, totals, int(iterations))

    def hist_labels_to_lut(self, keys: torch.Tensor, labels: torch.Tensor, lut: torch.Tensor):
        assert lut.numel() == self.HIST_BINS and lut.dtype == torch.uint8
        self._bind()
        self.ctx.call("llfe_hist_labels_to_lut", keys, labels, keys.numel(), lut)

    def pixels_lookup(self, bgr_rows: torch.Tensor, lut: torch.Tensor, labels: torch.Tensor):
        x = bgr_rows.contiguous()
        self._bind()
        self.ctx.call("llfe_pixels_lookup", x, x.numel() // 3, lut, labels)

    def kmeans_pixels_farthest(self, bgr_rows: torch.Tensor, centers: torch.Tensor, donor: int, base3,
                               index_base: int, out: torch.Tensor, skip=(), want_dist_bits: int = 0):
        """out[0] = max(out[0], code of the donor member farthest from base3); skip: global pixel
        indices to ignore (already moved by earlier repairs of the same update); want_dist_bits != 0: float32
        bits of the answer's distance when already known -- only pixels at exactly that distance are considered."""
        import ctypes

        x = bgr_rows.contiguous()
        arr = (ctypes.c_float * 3)(*[float(v) for v in base3])
        sk = (ctypes.c_uint32 * max(1, len(skip)))(*[int(v) for v in skip])
        self._bind()
        self.ctx.call("llfe_kmeans_pixels_farthest", x, x.numel() // 3, centers.shape[0], centers, int(donor),
                      ctypes.addressof(arr), int(index_base), ctypes.addressof(sk), len(skip), int(want_dist_bits), out)

    def kmeans_hist_farthest(self, keys: torch.Tensor, centers: torch.Tensor, donor: int, base3, out_bits: torch.Tensor):
        """out_bits[0] (int32) = max(out_bits[0], float32 bits of the largest distance to base3 among the colours
        of `keys` assigned to `donor`)."""
        import ctypes

        arr = (ctypes.c_float * 3)(*[float(v) for v in base3])
        self._bind()
        self.ctx.call("llfe_kmeans_hist_farthest", keys, keys.numel(), centers.shape[0], centers, int(donor),
                      ctypes.addressof(arr), out_bits)

    # -- fused pipeline -------------------------------------------------------------
    def pipeline(self, bgr: torch.Tensor, shapes: bool = True, shadows: bool = True, colors: bool = True,
                 noise: torch.Tensor | None = None, seed: int = 0, max_unique: int = 1 << 16, low: int = 50,
                 high: int = 150, out: dict | None = None) -> dict:
        """colours + shapes + shadows from one batch: returns a dict with
        shape_mask, shadow_mask, shadow_sums, keys, count (device tensors)."""
        x, _ = _batch(bgr, 3)
        n, h, w, _ = x.shape
        out = {} if out is None else out
        if shapes and "shape_mask" not in out:
            out["shape_mask"] = self._empty((n, h, w))
        if shadows and "shadow_mask" not in out:
            out["shadow_mask"] = self._empty((n, h, w))
            out["shadow_sums"] = self._empty((n, 2), torch.int64)
        if colors and "keys" not in out:
            out["keys"] = self._empty((n, max_unique), torch.int32)
            out["count"] = self._empty((n,), torch.int32)
        if noise is not None:
            noise = noise.contiguous()
        self._bind()
        self.ctx.call("llfe_pipeline", x, n, h, w, int(low), int(high), out.get("shape_mask") if shapes else None,
                      out.get("shadow_mask") if shadows else None, out.get("shadow_sums") if shadows else None,
                      noise, int(seed) & 0xFFFFFFFFFFFFFFFF, out.get("keys") if colors else None,
                      out.get("count") if colors else None, int(max_unique))
        return out
