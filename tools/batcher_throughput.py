#!/usr/bin/env python
"""Requests/s of the drop-in services one at a time vs through services.batching.RequestBatcher with C
concurrent clients (1080p synthetic design frames already decoded; contour tracing off in both arms)."""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from low_level_feature_extraction_b200.services import ColorExtractor, ShadowAnalyzer, ShapeAnalyzer
from low_level_feature_extraction_b200.services.batching import RequestBatcher
from low_level_feature_extraction_b200.synth import design_image

clients = int(sys.argv[1]) if len(sys.argv) > 1 else 32
per_client = int(sys.argv[2]) if len(sys.argv) > 2 else 16
imgs = [design_image(1080, 1920, s) for s in range(8)]
ColorExtractor.noise_mode = "device"

for _ in range(2):
    ColorExtractor.extract_colors(imgs[0]); ShapeAnalyzer.preprocess_image(imgs[0]); ShadowAnalyzer.analyze_shadow_level(imgs[0])
t0 = time.perf_counter()
n_single = 32
for i in range(n_single):
    im = imgs[i % 8]
    ColorExtractor.extract_colors(im)
    ShapeAnalyzer.preprocess_image(im)
    ShadowAnalyzer.analyze_shadow_level(im)
single = n_single / (time.perf_counter() - t0)

with RequestBatcher(device=0, max_batch=32, max_wait_ms=2.0, shapes=False) as rb:
    for _ in range(3):
        [f.result() for f in [rb.submit(imgs[i % 8]) for i in range(32)]]
    lat = []
    start = threading.Barrier(clients + 1)

    def client(c):
        start.wait()
        for j in range(per_client):
            t = time.perf_counter()
            rb.analyze(imgs[(c + j) % 8])
            lat.append(time.perf_counter() - t)

    th = [threading.Thread(target=client, args=(c,)) for c in range(clients)]
    for t in th:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    b0, n0 = rb.batches, rb.images
lat.sort()
print(json.dumps({"workload": "1080p frame -> colours + shape mask + shadow mask/level per request", "clients": clients,
                  "requests": clients * per_client, "one_at_a_time_rps": single, "batched_rps": clients * per_client / dt,
                  "batched_latency_ms_p50": lat[len(lat) // 2] * 1e3, "batched_latency_ms_p95": lat[int(len(lat) * 0.95)] * 1e3,
                  "mean_batch": n0 / max(1, b0)}))
