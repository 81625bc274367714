"""CPU timing of the reference's own path (bench.py's cpu_baseline / --impl reference legs).

TEST INFRASTRUCTURE.  Runs the reference implementation of the hot path on the
host cores: the UNMODIFIED reference classes when /root/reference is present
(build container), otherwise oracle/refpath.py -- the same call sequences against
the same OpenCV binary (kind "port").  One image per task over a process pool with
cv2.setNumThreads(1) in every worker, which is the CPU's best throughput for a
batch (SURVEY.md section 8(d) plan, variant (b)).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def _init(workload: str, h: int, w: int, use_reference: bool, k: int = 5):
    import cv2

    cv2.setNumThreads(1)
    from low_level_feature_extraction_b200.synth import design_image

    _STATE["workload"] = workload
    _STATE["k"] = k
    _STATE["imgs"] = [design_image(h, w, s) for s in range(2)]
    _STATE["ref"] = None
    if use_reference:
        from oracle import load_reference

        _STATE["ref"] = load_reference.load()


def _one(i: int) -> int:
    import cv2
    from oracle import refpath

    wl = _STATE["workload"]
    img = _STATE["imgs"][i % len(_STATE["imgs"])]
    ref = _STATE["ref"]
    np.random.seed(1000 + i)
    cv2.setRNGSeed(1000 + i)
    if wl in ("pipeline", "shapes"):
        (ref["ShapeAnalyzer"].preprocess_image if ref else refpath.shape_mask)(img)
    if wl in ("pipeline", "shadows", "palette_shadows"):
        (ref["ShadowAnalyzer"].analyze_shadow_level if ref else refpath.shadow_level)(img)
    if wl in ("pipeline", "colors", "palette_shadows"):
        (ref["ColorExtractor"].extract_colors if ref else refpath.extract_colors)(img, _STATE.get("k", 5))
    return i


class CpuReference:
    def __init__(self, workload: str, h: int, w: int, procs: int | None = None, k: int = 5):
        from oracle import load_reference

        self.workload, self.h, self.w = workload, h, w
        self.procs = procs or os.cpu_count() or 1
        self.kind = "reference" if load_reference.available() else "port"
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.procs, initializer=_init, initargs=(workload, h, w, self.kind == "reference", k))
        self.pool.map(_one, range(self.procs))  # warm: imports, synthetic inputs, OpenCV's lazy init

    def images_per_step(self, target_seconds: float = 8.0) -> int:
        t0 = time.perf_counter()
        self.pool.map(_one, range(self.procs))
        dt = max(time.perf_counter() - t0, 1e-3)
        rounds = max(1, min(16, int(target_seconds / dt)))
        return rounds * self.procs

    def step(self, n_images: int) -> float:
        """Process n_images; returns seconds."""
        t0 = time.perf_counter()
        self.pool.map(_one, range(n_images), chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
