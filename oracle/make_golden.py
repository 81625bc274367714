"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Inputs are seeded synthetic images (low_level_feature_extraction_b200.synth);
each fixture records the input's sha256 so the tests can prove they rebuilt
the same bytes, the reference's outputs (small arrays in full, large ones as
sha256 + a strided sample), and how each output was produced:
  "reference"  -- returned by the reference's own function, imported from
                  /root/reference (live .py or stale .pyc service);
  "cv2-callsite" -- the two text/font preprocess functions cannot be imported
                  here (their modules import pytesseract, not installed), so
                  their cv2 call sequence (text_extractor.py:27,40,43-44 and
                  font_detector.py:28,31-35) is executed directly.
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import os
import sys

import numpy as np
import cv2

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from low_level_feature_extraction_b200.synth import design_image, noise_image  # noqa: E402
from oracle import load_reference, refpath  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    return o


def main():
    ref = load_reference.load()
    os.makedirs(OUT, exist_ok=True)
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "cases": {}}

    # ---- masks + shapes + shadows -------------------------------------
    mask_cases = {
        "design_270x480_s1": design_image(270, 480, 1),
        "design_360x640_s2": design_image(360, 640, 2),
        "noise_96x160_s3": noise_image(96, 160, 3),
        "design_101x203_s4": design_image(101, 203, 4),   # odd sizes: unaligned rows
    }
    arrays = {}
    for name, img in mask_cases.items():
        SA, SH = ref["ShapeAnalyzer"], ref["ShadowAnalyzer"]
        smask = SA.preprocess_image(img)
        shapes = SA.analyze_shapes(img)
        blurred = SH.preprocess_image(img)
        level = SH.analyze_shadow_level(img)
        _, thr = refpath.shadow_mask(img)
        assert np.array_equal(blurred, refpath.shadow_mask(img)[0])
        arrays[name + "/shape_mask"] = smask
        arrays[name + "/shadow_blurred"] = blurred
        arrays[name + "/shadow_thresh"] = thr          # cv2-callsite (the reference returns only the level)
        arrays[name + "/text_mask"] = refpath.text_mask(img)   # cv2-callsite
        arrays[name + "/font_mask"] = refpath.font_mask(img)   # cv2-callsite
        meta["cases"][name] = {
            "input_sha256": sha(img), "shape": list(img.shape),
            "shadow_level": level, "analyze_shapes": jsonable(shapes),
            "provenance": {"shape_mask": "reference", "shadow_blurred": "reference", "shadow_level": "reference",
                           "analyze_shapes": "reference", "shadow_thresh": "cv2-callsite",
                           "text_mask": "cv2-callsite", "font_mask": "cv2-callsite"},
        }

    # ---- colours -------------------------------------------------------
    color_cases = [("design_270x480_s1", 5, 3), ("design_270x480_s1", 16, 11), ("design_360x640_s2", 5, 7),
                   ("noise_96x160_s3", 5, 5), ("design_101x203_s4", 8, 2)]
    meta["colors"] = []
    for name, k, seed in color_cases:
        img = mask_cases[name]
        np.random.seed(seed)
        cv2.setRNGSeed(seed)
        cf = ref["ColorExtractor"].extract_colors(img, k)
        # also the intermediate (centers, labels) of _get_dominant_colors under the same streams
        np.random.seed(seed)
        rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB).reshape(-1, 3)
        noise = np.random.normal(0, 0.5, rgb.shape).astype(np.int8)
        px = np.clip(rgb.astype(np.int32) + noise, 0, 255).astype(np.uint8)
        cv2.setRNGSeed(seed)
        centers, labels = ref["ColorExtractor"]._get_dominant_colors(px, k)
        tag = f"{name}/k{k}_seed{seed}"
        arrays[tag + "/centers"] = centers
        arrays[tag + "/labels"] = labels.astype(np.int32)
        meta["colors"].append({"case": name, "k": k, "seed": seed, "tag": tag,
                               "n_unique": int(len(np.unique(px, axis=0))),
                               "result": {"primary": cf.primary, "background": cf.background,
                                          "accent": list(cf.accent), "metadata": jsonable(cf.metadata)},
                               "provenance": "reference"})

    # ---- resize / preprocessing ------------------------------------------
    meta["resize"] = []
    big = design_image(600, 2400, 5)
    ok, png = cv2.imencode(".png", big)
    out = asyncio.run(ref["validate_and_preprocess_image"](png.tobytes(), "golden", "auto"))
    meta["resize"].append({"name": "auto_600x2400_s5", "kind": "validate_and_preprocess_image(auto)",
                           "in_shape": list(big.shape), "in_sha256": sha(big), "out_shape": list(out.shape),
                           "out_sha256": sha(out), "provenance": "reference"})
    arrays["auto_600x2400_s5/sample"] = out[::7, ::11].copy()
    big2 = design_image(1300, 2200, 6)
    out2 = ref["ImageProcessor"].resize_image(big2)
    meta["resize"].append({"name": "fit_1300x2200_s6", "kind": "ImageProcessor.resize_image(1920,1080)",
                           "in_shape": list(big2.shape), "in_sha256": sha(big2), "out_shape": list(out2.shape),
                           "out_sha256": sha(out2), "provenance": "reference"})
    arrays["fit_1300x2200_s6/sample"] = out2[::7, ::11].copy()
    src3 = design_image(512, 768, 7)
    IT = ref["ImageTransformer"]
    for nm, o in (("half", IT.resize(src3.copy(), width=384, height=256)),           # integer 2x path
                  ("quarter", IT.generate_thumbnail(src3.copy(), size=(192, 128))),   # integer 4x path
                  ("frac", IT.resize(src3.copy(), width=500, height=300))):           # fractional path
        meta["resize"].append({"name": f"it_{nm}_512x768_s7", "kind": f"ImageTransformer {nm}",
                               "in_shape": list(src3.shape), "in_sha256": sha(src3), "out_shape": list(o.shape),
                               "out_sha256": sha(o), "provenance": "reference"})
        arrays[f"it_{nm}_512x768_s7/out"] = o

    # ---- brightness/contrast + 3-channel blur --------------------------------
    src4 = design_image(120, 200, 8)
    meta["transform"] = {"input_sha256": sha(src4), "shape": list(src4.shape), "provenance": "reference"}
    arrays["transform/bc_1.2_0.9"] = IT.adjust_brightness_contrast(src4.copy(), brightness=1.2, contrast=0.9)
    arrays["transform/bc_1.3_1.5"] = IT.adjust_brightness_contrast(src4.copy(), brightness=1.3, contrast=1.5)
    arrays["transform/gaussian_blur"] = IT.apply_filter(src4.copy(), "gaussian_blur")

    np.savez_compressed(os.path.join(OUT, "golden.npz"), **{k.replace("/", "__"): v for k, v in arrays.items()})
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(arrays), "arrays;", os.path.getsize(os.path.join(OUT, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()
