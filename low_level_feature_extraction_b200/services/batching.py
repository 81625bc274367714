"""Micro-batching of concurrent single-image requests.

The reference serves one request at a time per worker process (`Dockerfile:26`); a single 1080p image
keeps a B200 busy for ~40 us of a ~1 ms call.  `RequestBatcher` sits between the request handlers and
the GPU: handlers `submit()` a decoded BGR frame and get a future; one worker thread collects whatever
arrived within `max_wait_ms` (up to `max_batch`), groups the frames by shape, and pushes every group
through ONE staged `BatchAnalyzer.run_host` call -- one read of each image produces the colour
palette, the shape mask and the shadow mask.  The per-image results carry what the three services
return: `ColorFeatures` (`ColorExtractor.extract_colors`), the `analyze_shapes` dict
(`ShapeAnalyzer.shapes_from_mask` on the GPU mask) and the shadow level string.

The palette uses device-generated noise (throughput mode); a caller that needs NumPy's
noise stream bit for bit uses `ColorExtractor.extract_colors` directly.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Callable

import numpy as np

from . import _runtime
from .color_extractor import ColorExtractor
from .shadow_analyzer import ShadowAnalyzer
from .shape_analyzer import ShapeAnalyzer

_STOP = object()


class RequestBatcher:
    def __init__(self, device: int = 0, max_batch: int = 32, max_wait_ms: float = 2.0, n_colors: int = 5,
                 shapes: bool = True, analyzer_factory: Callable[[int, int], Any] | None = None, host_threads: int = 8):
        """analyzer_factory(h, w) -> object with run_host(pinned (n, h, w, 3) uint8 tensor) -> dict of host
        tensors (`shape_mask`, `shadow_mask`, `shadow_sums`, `centers`, `k_used`, `cluster_sizes`); defaults to
        `BatchAnalyzer` on `device`.  shapes=False skips the host contour tracing (the mask is still returned).
        host_threads: the per-image host work of a batch -- copying the frame into the pinned staging buffer,
        copying the masks out, contour tracing -- is NumPy / OpenCV code that releases the GIL, so it is spread
        over a small thread pool (at 1080p the 10 MB of memcpy per image would otherwise dominate the batch)."""
        self.device = device
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self.n_colors = int(n_colors)
        self.shapes = shapes
        self._factory = analyzer_factory or self._default_factory
        self._analyzers: dict[tuple[int, int], Any] = {}
        self._staging: dict[tuple[int, int], Any] = {}
        self._host_out: dict[tuple[int, int], Any] = {}
        self._slot: dict[tuple[int, int], int] = {}
        self._q: queue.Queue = queue.Queue()
        self.batches = 0          # launches made
        self.images = 0           # images served
        self._closed = False
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(host_threads)), thread_name_prefix="llfe-batcher-host")
        self._worker = threading.Thread(target=self._loop, name="llfe-batcher", daemon=True)
        self._worker.start()

    # ---- client side ---------------------------------------------------------------------------
    def submit(self, image: np.ndarray) -> Future:
        """image: (H, W, 3) uint8 BGR.  The future resolves to {'colors': ColorFeatures, 'shapes': dict | None,
        'shadow_level': str, 'shape_mask': (H, W) u8, 'shadow_mask': (H, W) u8}."""
        if self._closed:
            raise RuntimeError("RequestBatcher is closed")
        fut: Future = Future()
        try:
            img = _runtime.as_bgr_u8(image)
        except Exception as e:   # bad input fails its own request, never the batch
            fut.set_exception(e)
            return fut
        self._q.put((img, fut))
        return fut

    def analyze(self, image: np.ndarray) -> dict:
        return self.submit(image).result()

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(_STOP)
            self._worker.join()
            self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- worker ----------------------------------------------------------------------------------
    def _default_factory(self, h: int, w: int):
        from ..batch import BatchAnalyzer, BatchConfig

        return BatchAnalyzer(self.device, h, w, BatchConfig(k=self.n_colors, host_chunk=min(16, self.max_batch),
                                                            contours=self.shapes))

    def _loop(self) -> None:
        # Two batches overlap: while the per-image host work of batch i (mask copies, contour geometry, palettes) runs,
        # batch i + 1 is already being copied in and computed (BatchAnalyzer.run_host_async).  A batch in flight is
        # finished as soon as no new request is waiting, so a lone client never waits for a follower.
        stop = False
        pending = None
        while not stop:
            if pending is not None:
                try:
                    item = self._q.get_nowait()
                except queue.Empty:
                    self._complete(pending)
                    pending = None
                    continue
            else:
                item = self._q.get()
            if item is _STOP:
                break
            batch = [item]
            deadline = time.perf_counter() + self.max_wait
            while len(batch) < self.max_batch:
                try:
                    nxt = self._q.get(timeout=max(0.0, deadline - time.perf_counter()))
                except queue.Empty:
                    break
                if nxt is _STOP:
                    stop = True
                    break
                batch.append(nxt)
            groups: dict[tuple[int, int], list] = {}
            for img, fut in batch:
                groups.setdefault(img.shape[:2], []).append((img, fut))
            for shape, items in groups.items():
                started = self._start(shape, items)
                if pending is not None:
                    self._complete(pending)
                pending = started
        if pending is not None:
            self._complete(pending)

    def _start(self, shape: tuple[int, int], items: list):
        """Stage the frames and enqueue the batch; returns what `_complete` needs (None if the batch already failed)."""
        import torch

        h, w = shape
        futs = [f for _, f in items]
        try:
            an = self._analyzers.get(shape)
            if an is None:
                an = self._analyzers[shape] = self._factory(h, w)
                self._staging[shape] = []
                self._host_out[shape] = []
                for _ in range(2):          # one set of pinned buffers per batch in flight
                    st = torch.empty((self.max_batch, h, w, 3), dtype=torch.uint8)
                    self._staging[shape].append(st.pin_memory() if torch.cuda.is_available() else st)
                    if hasattr(an, "alloc_host_outputs"):
                        self._host_out[shape].append(an.alloc_host_outputs(self.max_batch))
                self._slot[shape] = 0
            slot = self._slot[shape]
            self._slot[shape] = slot ^ 1
            stage = self._staging[shape][slot]
            m = len(items)
            view = stage.numpy()
            list(self._pool.map(lambda i: np.copyto(view[i], items[i][0]), range(m)))
            host_out = self._host_out[shape][slot] if self._host_out[shape] else None
            with _runtime.lock():
                if hasattr(an, "run_host_async"):
                    call = an.run_host_async(stage[:m], host_out) if host_out is not None else an.run_host_async(stage[:m])
                    return (shape, items, an, m, call, None)
                full = an.run_host(stage[:m], host_out) if host_out is not None else an.run_host(stage[:m])
                return (shape, items, an, m, None, full)
        except Exception as e:
            for fut in futs:
                if not fut.done():
                    fut.set_exception(e)
            return None

    def _complete(self, started) -> None:
        if started is None:
            return
        shape, items, an, m, call, full = started
        h, w = shape
        futs = [f for _, f in items]
        try:
            if call is not None:
                with _runtime.lock():
                    full = call.result()
            out = {k: (v[:m] if hasattr(v, "shape") else v) for k, v in full.items()}
            self.batches += 1
            self.images += m
            centers = out["centers"].numpy()
            k_used = out["k_used"].numpy()
            sizes = out["cluster_sizes"].numpy()
            sums = out["shadow_sums"].numpy()
            shape_masks = out["shape_mask"].numpy()
            shadow_masks = out["shadow_mask"].numpy()

            def finish(i: int) -> None:
                fut = futs[i]
                try:
                    k = int(k_used[i])
                    if k < 0:
                        raise RuntimeError("unique-colour list truncated and not resolved by the analyzer")
                    if self.n_colors <= 1:
                        # color_extractor.py:185-186 returns the whole unique list as "centres" in this case: take
                        # the single-image path, which brings the list to the host (device noise, as the batch)
                        c8, lab = ColorExtractor._dominant_from_bgr(items[i][0], None, self.n_colors)
                        colors = ColorExtractor._palette_from_clusters(
                            c8, np.bincount(lab, minlength=len(c8)) if len(c8) > 1 else None)
                    else:
                        c8 = centers[i, :k].astype(np.uint8)                  # color_extractor.py:197 truncation
                        colors = ColorExtractor._palette_from_clusters(c8, sizes[i, :k].astype(np.int64) if k > 1 else None)
                    shape_mask = shape_masks[i].copy()
                    shapes = None
                    if self.shapes:
                        conts = an.contours(out, i) if "contour_counts" in out else None
                        shapes = (ShapeAnalyzer.shapes_from_contours(conts, w, h) if conts is not None
                                  else ShapeAnalyzer.shapes_from_mask(shape_mask, w, h))
                    res = {"colors": colors,
                           "shapes": shapes,
                           "shadow_level": ShadowAnalyzer.level_from_sums(int(sums[i, 0]), int(sums[i, 1])),
                           "shape_mask": shape_mask,
                           "shadow_mask": shadow_masks[i].copy()}
                    fut.set_result(res)
                except Exception as e:
                    fut.set_exception(e)

            list(self._pool.map(finish, range(m)))   # the pinned buffers are reused by the next batch: wait
        except Exception as e:
            for fut in futs:
                if not fut.done():
                    fut.set_exception(e)
