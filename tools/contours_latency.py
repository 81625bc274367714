#!/usr/bin/env python
"""Latency of the device contour finder (csrc/k_contours.cu) on 1080p masks, kernel by kernel (CUDA events of the
library's own profile hooks) and end to end, next to cv2.findContours on the host.

    python tools/contours_latency.py [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2
import numpy as np
import torch


def med(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import low_level_feature_extraction_b200 as pkg
    from low_level_feature_extraction_b200 import contours as ct
    from low_level_feature_extraction_b200.services import ShapeAnalyzer, _runtime
    from low_level_feature_extraction_b200.synth import design_image
    from oracle import refpath

    eng = pkg.engine(0)
    img = design_image(1080, 1920, 0)
    out = {"unit": "ms (median)", "reps": a.reps}
    for name, mask, min_area2 in (("shape mask (dilated Canny edges)", refpath.shape_mask(img), 200),
                                  ("font mask (adaptive threshold)", refpath.font_mask(img), 0)):
        d = torch.from_numpy(mask).cuda()
        rec = {"external_contours": len(cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0])}
        eng.contours_external(d, min_area2, 1 << 15, 1 << 19)
        torch.cuda.synchronize()
        eng.ctx.profile_begin()
        for _ in range(a.reps):
            eng.contours_external(d, min_area2, 1 << 15, 1 << 19)
        torch.cuda.synchronize()
        prof = eng.ctx.profile_end()
        rec["kernels_ms"] = {k: round(v["ms"] / a.reps, 5) for k, v in prof.items()}

        def dev():
            eng.contours_external(d, min_area2, 1 << 15, 1 << 19)
            torch.cuda.synchronize()

        rec["device_call_ms"] = med(dev, a.reps)
        rec["host_mask_call_ms"] = med(lambda: ct.find_external_host(_runtime.context(), mask, min_area2), a.reps)
        rec["cv2_findContours_ms"] = med(lambda: cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE), a.reps)
        out[name] = rec
    out["ShapeAnalyzer.analyze_shapes_ms"] = med(lambda: ShapeAnalyzer.analyze_shapes(img), a.reps)
    out["ShapeAnalyzer.preprocess_image_ms"] = med(lambda: ShapeAnalyzer.preprocess_image(img), a.reps)
    conts = ShapeAnalyzer._device_contours(img)
    out["host_tail_shapes_from_contours_ms"] = med(lambda: ShapeAnalyzer.shapes_from_contours(conts, 1920, 1080), a.reps)
    out["reference_cpu_port_analyze_shapes_ms"] = med(lambda: refpath.analyze_shapes_from_mask(refpath.shape_mask(img), img.shape), a.reps)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
