"""Supplementary golden fixtures: the reference's `performance` (INTER_LINEAR) and `high_quality`
(INTER_LANCZOS4) preprocessing modes (app/services/analyze/utils.py:128-143).

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden_resize_modes

Runs the UNMODIFIED `validate_and_preprocess_image` on PNG encodings of seeded synthetic images and
writes tests/golden/golden_resize_modes.json (input sha256, output shape, output sha256, and a strided
sample of the output so a mismatch can be localised).
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from low_level_feature_extraction_b200.synth import design_image  # noqa: E402
from oracle import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_resize_modes.json")

CASES = [  # (name, h, w, seed, mode)
    ("performance_1080x1920_s21", 1080, 1920, 21, "performance"),
    ("performance_700x1300_s22", 700, 1300, 22, "performance"),
    ("performance_2000x1000_s23", 2000, 1000, 23, "performance"),        # exactly 2x: OpenCV reroutes to INTER_AREA
    ("high_quality_1100x4400_s24", 1100, 4400, 24, "high_quality"),
    ("high_quality_4100x820_s25", 4100, 820, 25, "high_quality"),
    ("high_quality_900x1600_s26", 900, 1600, 26, "high_quality"),        # below the limit: unchanged
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = load_reference.load()
    fn = ref["validate_and_preprocess_image"]
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "producer": "reference validate_and_preprocess_image", "cases": {}}
    for name, h, w, seed, mode in CASES:
        img = design_image(h, w, seed)
        ok, png = cv2.imencode(".png", img)
        assert ok
        out = asyncio.run(fn(png.tobytes(), "golden", mode))
        meta["cases"][name] = {"h": h, "w": w, "seed": seed, "mode": mode, "input_sha256": sha(img),
                               "out_shape": list(out.shape), "out_sha256": sha(out),
                               "sample": out[::97, ::89].reshape(-1)[:64].astype(int).tolist()}
        print(name, out.shape)
    with open(OUT, "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
