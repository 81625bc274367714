#!/usr/bin/env python
"""BASELINE config 1 and friends: single-image latency of the drop-in service calls on one GPU next to the
reference's own call sequence (oracle/refpath.py: the same cv2 / numpy calls) on the host cores.

    python tools/service_latency.py [--height 1080 --width 1920 --reps 5]

Prints one JSON line with median milliseconds per call.  The colour service is timed in both noise modes:
"numpy" (bit-compatible with the reference under np.random.seed; the 6.2 M-sample Gaussian draw on the host
dominates) and "device" (noise of the same distribution generated in the kernel).
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2
import numpy as np


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
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from low_level_feature_extraction_b200.services import (ColorExtractor, ShadowAnalyzer, ShapeAnalyzer,
                                                            validate_and_preprocess_image)
    from low_level_feature_extraction_b200.synth import design_image
    from oracle import refpath  # CPU side of the comparison

    img = design_image(a.height, a.width, 0)
    ok, png = cv2.imencode(".png", img)
    png = png.tobytes()
    out = {"image": f"{a.width}x{a.height} synthetic design image, PNG {len(png)} bytes", "unit": "ms (median)",
           "host_threads_cv2": cv2.getNumThreads()}
    ours, ref = {}, {}
    ours["validate_and_preprocess_image(auto)"] = med(lambda: asyncio.run(validate_and_preprocess_image(png, "t", "auto")), a.reps)
    ref["validate_and_preprocess_image(auto)"] = med(lambda: refpath.auto_resize(cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_COLOR)), a.reps)
    ours["ShapeAnalyzer.analyze_shapes"] = med(lambda: ShapeAnalyzer.analyze_shapes(img), a.reps)
    ref["ShapeAnalyzer.analyze_shapes"] = med(lambda: refpath.analyze_shapes_from_mask(refpath.shape_mask(img), img.shape), a.reps)
    ours["ShadowAnalyzer.analyze_shadow_level"] = med(lambda: ShadowAnalyzer.analyze_shadow_level(img), a.reps)
    ref["ShadowAnalyzer.analyze_shadow_level"] = med(lambda: refpath.shadow_level(img), a.reps)
    ColorExtractor.noise_mode = "numpy"
    ours["ColorExtractor.extract_colors(k=5) numpy noise"] = med(lambda: ColorExtractor.extract_colors(img, 5), a.reps)
    ColorExtractor.noise_mode = "device"
    ours["ColorExtractor.extract_colors(k=5) device noise"] = med(lambda: ColorExtractor.extract_colors(img, 5), a.reps)
    ref["ColorExtractor.extract_colors(k=5)"] = med(lambda: refpath.extract_colors(img, 5), max(2, a.reps // 2))
    out["ours_gpu"] = ours
    out["reference_cpu_port"] = ref
    c1 = ours["validate_and_preprocess_image(auto)"] + ours["ColorExtractor.extract_colors(k=5) device noise"]
    r1 = ref["validate_and_preprocess_image(auto)"] + ref["ColorExtractor.extract_colors(k=5)"]
    out["config1_images_per_sec"] = {"ours_device_noise": 1e3 / c1, "reference_cpu_port": 1e3 / r1}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
