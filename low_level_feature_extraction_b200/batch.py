"""Batched colours + shapes + shadows over many images (BASELINE configs 2-4).

`BatchAnalyzer.run_device` works on a device-resident (B, H, W, 3) uint8 BGR batch
and leaves every result on the device.  `run_host` is the end-to-end path: pinned
host images in, host results out, with the host<->device copies of one chunk
overlapping the kernels of the others (`host_streams` streams, one llfe context each
because a context's workspace belongs to one stream at a time).  PCIe is the bound:
1.59 GB in + 1.06 GB out per 256 images at ~56 GB/s per direction.

Batches shard across GPUs by image index with no collective (see dist.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .ops import Engine


@dataclass
class BatchConfig:
    colors: bool = True
    shapes: bool = True
    shadows: bool = True
    k: int = 5                 # ColorExtractor default n_colors
    attempts: int = 10         # cv2.kmeans attempts of the reference
    max_iter: int = 200
    eps: float = 0.2
    # capacity of the per-image unique-colour list of the batched launch.  An image with more colours (any
    # photo-like frame) is never clustered on a truncated list: the batched k-means skips it (k_used = -1)
    # and `resolve_overflow` redoes it alone with a list sized from its own count.
    max_unique: int = 1 << 16
    low: int = 50
    high: int = 150
    seed: int = 0              # device noise seed / cv::RNG state base
    chunk: int = 256           # images per C-ABI call on the device path (the library sub-chunks for L2)
    host_chunk: int = 16       # images per copy/compute pipeline stage of run_host
    host_streams: int = 3      # pipeline depth of run_host (streams / llfe contexts / staging buffers)
    # run_host: masks cross PCIe as bit planes (P/8 bytes instead of P) and are expanded to the reference's u8 arrays
    # by host threads while later stages are copied and computed (the copy back otherwise costs the inbound direction
    # ~8 % on one GPU and more than half when eight ranks share the host's PCIe fabric)
    packed_masks: bool = True
    expand_threads: int = 4    # host threads per expansion call
    expand_workers: int = 3    # stages being expanded at the same time
    # external contours of the shape mask on the device (cv2.findContours(EXTERNAL, SIMPLE) + the reference's
    # `contourArea < 100` filter): per image `max_contours` header records and `max_points` vertices; an image that
    # needs more reports it in contour_counts (column 0 > max_contours or column 2 != 0) and the caller redoes it alone
    contours: bool = False
    max_contours: int = 1024
    max_points: int = 8192


class HostCall:
    """One batch of `BatchAnalyzer.run_host_async` in flight."""

    def __init__(self, owner, images, host_out, n, done, pending, bytes_in, bytes_out, slot):
        self.owner, self.images, self.host_out, self.n = owner, images, host_out, n
        self.done, self.pending, self.bytes_in, self.bytes_out, self.slot = done, pending, bytes_in, bytes_out, slot
        self._result = None

    def result(self) -> dict:
        if self._result is None:
            self._result = self.owner._finish_host(self)
        return self._result


class BatchAnalyzer:
    def __init__(self, device: int, h: int, w: int, cfg: BatchConfig | None = None):
        self.cfg = cfg or BatchConfig()
        self.h, self.w = h, w
        self.device = torch.device("cuda", device)
        ns = max(1, self.cfg.host_streams)
        self.engines = [Engine(device) for _ in range(ns)]
        self.streams = [torch.cuda.Stream(self.device) for _ in range(ns)]
        self._dev_in = None
        self._dev_bits = None
        self._host_bits = [None, None]     # pinned bit planes, one set per batch in flight
        self._inflight = [None, None]
        self._calls = 0
        self._pool = None

    # ---- device-resident ---------------------------------------------------------------
    def alloc_outputs(self, n: int) -> dict:
        c, d = self.cfg, self.device
        out = {}
        if c.shapes:
            out["shape_mask"] = torch.empty((n, self.h, self.w), dtype=torch.uint8, device=d)
        if c.shapes and c.contours:
            out["contour_headers"] = torch.empty((n, c.max_contours, 10), dtype=torch.int32, device=d)
            out["contour_points"] = torch.empty((n, c.max_points, 2), dtype=torch.int32, device=d)
            out["contour_counts"] = torch.empty((n, 4), dtype=torch.int32, device=d)
        if c.shadows:
            out["shadow_mask"] = torch.empty((n, self.h, self.w), dtype=torch.uint8, device=d)
            out["shadow_sums"] = torch.empty((n, 2), dtype=torch.int64, device=d)
        if c.colors:
            out["keys"] = torch.empty((n, c.max_unique), dtype=torch.int32, device=d)
            out["count"] = torch.empty((n,), dtype=torch.int32, device=d)
            out["centers"] = torch.zeros((n, c.k, 3), dtype=torch.float32, device=d)
            out["labels"] = torch.empty((n, c.max_unique), dtype=torch.int32, device=d)
            out["k_used"] = torch.empty((n,), dtype=torch.int32, device=d)
            out["cluster_sizes"] = torch.empty((n, c.k), dtype=torch.int32, device=d)
            out["status"] = torch.empty((n,), dtype=torch.int32, device=d)
            out["rng"] = (torch.arange(n, dtype=torch.int64, device=d) + 1000 + c.seed)
        return out

    def run_device(self, bgr: torch.Tensor, out: dict | None = None, engine: Engine | None = None,
                   noise: torch.Tensor | None = None, resolve: bool = True) -> dict:
        """All results stay on the device.  With resolve=True (default) the call ends by reading the unique-colour
        counts (one small device->host copy, i.e. it synchronises torch's current stream) and redoing every image
        whose list did not fit `max_unique`; resolve=False stays asynchronous and leaves k_used = -1 for those
        images (call `resolve_overflow` before using the palettes)."""
        c = self.cfg
        eng = engine or self.engines[0]
        n = bgr.shape[0]
        out = out if out is not None else self.alloc_outputs(n)
        for i0 in range(0, n, c.chunk):
            sl = slice(i0, min(n, i0 + c.chunk))
            view = {k: v[sl] for k, v in out.items()}
            if c.colors:
                # masks + unique colours + k-means in one call: its two chains run on two streams (llfe_analyze)
                x = bgr[sl].contiguous()
                nz = None if noise is None else noise[sl].contiguous()
                eng._bind()
                eng.ctx.call("llfe_analyze", x, sl.stop - sl.start, self.h, self.w, int(c.low), int(c.high),
                             view["shape_mask"] if c.shapes else None, view["shadow_mask"] if c.shadows else None,
                             view["shadow_sums"] if c.shadows else None, nz, (c.seed + i0) & 0xFFFFFFFFFFFFFFFF,
                             view["keys"], view["count"], c.max_unique, c.k, c.attempts, c.max_iter, float(c.eps),
                             view["rng"], view["centers"], view["labels"], view["k_used"], view["cluster_sizes"],
                             view["status"])
            else:
                eng.pipeline(bgr[sl], shapes=c.shapes, shadows=c.shadows, colors=False, low=c.low, high=c.high, out=view)
            if c.shapes and c.contours:
                eng._bind()
                eng.ctx.call("llfe_contours_external", view["shape_mask"], sl.stop - sl.start, self.h, self.w, 200,
                             view["contour_headers"], c.max_contours, view["contour_points"], c.max_points,
                             view["contour_counts"])
        if c.colors and resolve:
            self.resolve_overflow(bgr, out, eng, noise)
        return out

    def resolve_overflow(self, bgr: torch.Tensor, out: dict, engine: Engine | None = None,
                         noise: torch.Tensor | None = None, counts=None) -> list:
        """Redo the palette of every image whose unique-colour list overflowed `max_unique` (count > max_unique,
        k_used = -1), each with a list sized from its own count and the SAME noise as the batched pass.
        centers / k_used / cluster_sizes / status of those images are overwritten; their keys / labels rows keep the
        truncated list.  Returns the indices that were redone.  counts: the host copy of out["count"] if the
        caller already has it."""
        c = self.cfg
        eng = engine or self.engines[0]
        cnt = out["count"].cpu() if counts is None else counts
        redo = [int(i) for i in torch.nonzero(cnt > c.max_unique).flatten()]
        for i in redo:
            i0 = (i // c.chunk) * c.chunk          # run_device seeds every chunk with c.seed + i0
            cen, kused, sizes, status = eng.palette_large(
                bgr[i], int(cnt[i]), c.k, 1000 + c.seed + i, seed=c.seed + i0, first_image=i - i0,
                noise=None if noise is None else noise[i], attempts=c.attempts, max_iter=c.max_iter, eps=c.eps)
            out["centers"][i].copy_(cen)
            out["k_used"][i:i + 1].copy_(kused)
            out["cluster_sizes"][i].copy_(sizes)
            out["status"][i:i + 1].copy_(status)
        return redo

    # ---- end to end: pinned host in, host out -------------------------------------------------
    def alloc_host_outputs(self, n: int) -> dict:
        c = self.cfg
        out = {}
        if c.shapes:
            out["shape_mask"] = torch.empty((n, self.h, self.w), dtype=torch.uint8).pin_memory()
        if c.shapes and c.contours:
            out["contour_headers"] = torch.empty((n, c.max_contours, 10), dtype=torch.int32).pin_memory()
            out["contour_points"] = torch.empty((n, c.max_points, 2), dtype=torch.int32).pin_memory()
            out["contour_counts"] = torch.empty((n, 4), dtype=torch.int32).pin_memory()
        if c.shadows:
            out["shadow_mask"] = torch.empty((n, self.h, self.w), dtype=torch.uint8).pin_memory()
            out["shadow_sums"] = torch.empty((n, 2), dtype=torch.int64).pin_memory()
        if c.colors:
            out["centers"] = torch.empty((n, c.k, 3), dtype=torch.float32).pin_memory()
            out["count"] = torch.empty((n,), dtype=torch.int32).pin_memory()
            out["k_used"] = torch.empty((n,), dtype=torch.int32).pin_memory()
            out["cluster_sizes"] = torch.empty((n, c.k), dtype=torch.int32).pin_memory()
            out["status"] = torch.empty((n,), dtype=torch.int32).pin_memory()
        return out

    def contours(self, host_out: dict, i: int):
        """cv2-ordered contours (area >= 100) of image i of a `run_host` result, or None when the image needs more
        than max_contours / max_points (the caller then takes the single-image path)."""
        from . import contours as ct

        cnt = host_out["contour_counts"][i].numpy()
        if cnt[0] > self.cfg.max_contours or cnt[2]:
            return None
        headers = host_out["contour_headers"][i, :int(cnt[0])].numpy().view(ct.HEADER).reshape(-1)
        return ct.to_cv2_contours(headers, host_out["contour_points"][i].numpy())

    def palettes(self, host_out: dict) -> list:
        """The reference's palette tail (color_extractor.py:231-284: bincount order, hex, white / black filter, primary /
        accents / background) for every image of a `run_host` result -> list of ColorFeatures."""
        from .services.color_extractor import ColorExtractor

        k_used = host_out["k_used"].numpy()
        if (k_used < 0).any():
            raise RuntimeError("unique-colour list truncated and not resolved")
        return ColorExtractor._palettes_from_batch(host_out["centers"].numpy(), k_used, host_out["cluster_sizes"].numpy())

    def _stages(self, n: int):
        """(first image, count) of every pipeline stage: short stages at both ends so that the un-overlapped
        first upload and last download are small, full `host_chunk` stages in between."""
        c = self.cfg.host_chunk
        ramp = []
        s = max(1, c // 4)
        while s < c:
            ramp.append(s)
            s *= 2
        sizes, left = [], n
        for s in ramp:                      # ramp up
            if left - s < sum(ramp):        # keep room for the ramp down
                break
            sizes.append(s)
            left -= s
        tail = [s for s in reversed(ramp)]
        while left > sum(tail):
            m = min(c, left - sum(tail))
            sizes.append(m)
            left -= m
        for s in tail:
            m = min(s, left)
            if m > 0:
                sizes.append(m)
                left -= m
        if left > 0:
            sizes.append(left)
        out, i0 = [], 0
        for m in sizes:
            out.append((i0, m))
            i0 += m
        assert i0 == n
        return out

    def run_host(self, images: torch.Tensor, host_out: dict | None = None) -> dict:
        """images: pinned CPU uint8 (B, H, W, 3).  Returns host tensors; synchronous."""
        return self.run_host_async(images, host_out).result()

    def run_host_async(self, images: torch.Tensor, host_out: dict | None = None) -> "HostCall":
        """Enqueue one batch (copies in, kernels, copies out, mask expansion on host threads) and return a handle;
        `handle.result()` waits for this batch only and returns the host tensors.  Up to two batches may be in flight:
        the first stages of batch i + 1 are copied in while the last stage of batch i is still being computed and copied
        back, which a synchronous call per batch cannot hide.  `images` and `host_out` must stay untouched until
        `result()` returns."""
        c = self.cfg
        n = images.shape[0]
        host_out = host_out if host_out is not None else self.alloc_host_outputs(n)
        mask_keys = [k for k in ("shape_mask", "shadow_mask") if k in host_out]
        packed = c.packed_masks and len(mask_keys) > 0
        # the batch that used this slot's pinned bit planes two calls ago must be done with them
        slot = self._calls % 2
        self._calls += 1
        if self._inflight[slot] is not None:
            self._inflight[slot].result()
        if self._dev_in is None:
            ns = len(self.streams)
            self._dev_in = [torch.empty((c.host_chunk, self.h, self.w, 3), dtype=torch.uint8, device=self.device) for _ in range(ns)]
            self._dev_out = [self.alloc_outputs(c.host_chunk) for _ in range(ns)]
        if packed and (self._host_bits[slot] is None or self._host_bits[slot].shape[1] < n):
            from concurrent.futures import ThreadPoolExecutor

            wpr = self.engines[0].ctx.lib.llfe_mask_bits_words_per_row(self.w)
            if self._dev_bits is None:
                self._dev_bits = [torch.empty((len(mask_keys), c.host_chunk, self.h, wpr), dtype=torch.int32, device=self.device)
                                  for _ in range(len(self.streams))]
            self._host_bits[slot] = torch.empty((len(mask_keys), n, self.h, wpr), dtype=torch.int32).pin_memory()
            if self._pool is None:
                self._pool = ThreadPoolExecutor(max(1, c.expand_workers), thread_name_prefix="llfe-expand")
        lib = self.engines[0].ctx.lib
        host_bits = self._host_bits[slot]
        pending = []

        def expand(ev, i0, m):
            ev.synchronize()                     # releases the GIL; the expansion below is a ctypes call (ditto)
            for j, key in enumerate(mask_keys):
                rc = lib.llfe_expand_mask_bits_host(host_bits[j, i0:i0 + m].data_ptr(), m, self.h, self.w,
                                                    host_out[key][i0:i0 + m].data_ptr(), c.expand_threads)
                if rc != 0:
                    raise RuntimeError(lib.llfe_last_error().decode())

        bytes_in = bytes_out = 0
        for j, (i0, m) in enumerate(self._stages(n)):
            b = j % len(self.streams)
            st = self.streams[b]
            with torch.cuda.stream(st):
                din = self._dev_in[b][:m]
                din.copy_(images[i0:i0 + m], non_blocking=True)
                bytes_in += din.numel()
                dout = {k: v[:m] for k, v in self._dev_out[b].items()}
                self.run_device(din, dout, engine=self.engines[b], resolve=False)
                small = ["shadow_sums", "centers", "count", "k_used", "cluster_sizes", "status",
                         "contour_headers", "contour_points", "contour_counts"]
                if packed:
                    eng = self.engines[b]
                    for q, key in enumerate(mask_keys):
                        bits = self._dev_bits[b][q, :m]
                        eng.ctx.call("llfe_pack_mask_bits", dout[key], m, self.h, self.w, bits)
                        host_bits[q, i0:i0 + m].copy_(bits, non_blocking=True)
                        bytes_out += bits.numel() * 4
                    ev = torch.cuda.Event()
                    ev.record(st)
                    pending.append(self._pool.submit(expand, ev, i0, m))
                else:
                    small = mask_keys + small
                for key in small:
                    if key in host_out:
                        host_out[key][i0:i0 + m].copy_(dout[key], non_blocking=True)
                        bytes_out += dout[key].numel() * dout[key].element_size()
        done = []
        for st in self.streams:           # this batch's work only: a later batch may already be behind it on the streams
            ev = torch.cuda.Event()
            ev.record(st)
            done.append(ev)
        call = HostCall(self, images, host_out, n, done, pending, bytes_in, bytes_out, slot)
        self._inflight[slot] = call
        return call

    def _finish_host(self, call: "HostCall") -> dict:
        c = self.cfg
        images, host_out, n = call.images, call.host_out, call.n
        bytes_in, bytes_out = call.bytes_in, call.bytes_out
        for ev in call.done:
            ev.synchronize()
        for f in call.pending:
            f.result()
        if c.colors:
            # images whose colour list overflowed the batched capacity: upload again, redo alone (rare: photo-like
            # frames), patch the host results.  The stage's chunk-local index fixes the noise.
            over = [int(i) for i in torch.nonzero(host_out["count"][:n] > c.max_unique).flatten()]
            stage_of = {}
            for (i0, m) in self._stages(n):
                for i in range(i0, i0 + m):
                    stage_of[i] = i0
            eng = self.engines[0]
            for i in over:
                with torch.cuda.stream(self.streams[0]):
                    din = self._dev_in[0][:1]
                    din.copy_(images[i:i + 1], non_blocking=True)
                    bytes_in += din.numel()
                    # run_device saw this image at index i - i0 of a call seeded with c.seed (stage-local chunk 0)
                    cen, kused, sizes, status = eng.palette_large(
                        din[0], int(host_out["count"][i]), c.k, 1000 + c.seed + (i - stage_of[i]), seed=c.seed,
                        first_image=i - stage_of[i], attempts=c.attempts, max_iter=c.max_iter, eps=c.eps)
                    host_out["centers"][i].copy_(cen, non_blocking=True)
                    host_out["k_used"][i:i + 1].copy_(kused, non_blocking=True)
                    host_out["cluster_sizes"][i].copy_(sizes, non_blocking=True)
                    host_out["status"][i:i + 1].copy_(status, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.streams[0])
                ev.synchronize()
        host_out["_h2d_bytes"] = bytes_in
        host_out["_d2h_bytes"] = bytes_out
        if self._inflight[call.slot] is call:
            self._inflight[call.slot] = None
        return host_out
