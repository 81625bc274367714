"""The drop-in service classes (host buffers in, reference-shaped results out) against the
outputs of the unmodified reference captured in tests/golden/."""
import asyncio
import hashlib
import os

import cv2
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cvops, refpath  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image  # noqa: E402
from low_level_feature_extraction_b200.services import (ColorExtractor, FontDetector, ImageProcessor,  # noqa: E402
                                                        ImageTransformer, ShadowAnalyzer, ShapeAnalyzer,
                                                        TextExtractor, validate_and_preprocess_image)

CASES = ["design_270x480_s1", "design_360x640_s2", "noise_96x160_s3", "design_101x203_s4"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_shapes_and_shadows_json(golden, golden_inputs, name):
    meta, arrays = golden
    img = golden_inputs[name]
    assert np.array_equal(ShapeAnalyzer.preprocess_image(img), arrays[name + "/shape_mask"])
    got = ShapeAnalyzer.analyze_shapes(img)
    want = meta["cases"][name]["analyze_shapes"]
    assert got["total_shapes"] == want["total_shapes"] and got["metadata"] == want["metadata"]
    for a, b in zip(got["shapes"], want["shapes"]):
        assert a == b
    inst = ShapeAnalyzer().extract_shapes(img)
    assert [s["type"] for s in inst] == [s["type"] for s in want["shapes"]]
    assert np.array_equal(ShadowAnalyzer.preprocess_image(img), arrays[name + "/shadow_blurred"])
    assert ShadowAnalyzer.analyze_shadow_level(img) == meta["cases"][name]["shadow_level"]
    assert np.array_equal(FontDetector.preprocess_image(img), arrays[name + "/font_mask"])
    assert np.array_equal(TextExtractor.preprocess_image(img), arrays[name + "/text_mask"])


def test_extract_colors_json_matches_reference(golden, golden_inputs):
    meta, _ = golden
    for c in meta["colors"]:
        img = golden_inputs[c["case"]]
        np.random.seed(c["seed"])                  # what a caller of the reference does ...
        ColorExtractor.set_rng_seed(c["seed"])     # ... plus the stand-in for cv2.setRNGSeed
        cf = ColorExtractor.extract_colors(img, c["k"])
        assert cf.primary == c["result"]["primary"], c["tag"]
        assert cf.background == c["result"]["background"], c["tag"]
        assert list(cf.accent) == c["result"]["accent"], c["tag"]
        assert cf.metadata == c["result"]["metadata"], c["tag"]


def test_extract_colors_rng_state_carries_like_cv2():
    """Two consecutive calls under one seeding == the reference's two calls (global RNG state carries over)."""
    img = design_image(90, 120, 5)
    np.random.seed(4)
    cv2.setRNGSeed(4)
    want = [refpath.extract_colors(img, 5), refpath.extract_colors(img, 4)]
    np.random.seed(4)
    ColorExtractor.set_rng_seed(4)
    got = [ColorExtractor.extract_colors(img, 5), ColorExtractor.extract_colors(img, 4)]
    for g, w in zip(got, want):
        assert {"primary": g.primary, "background": g.background, "accent": list(g.accent)} == w


def test_extract_colors_edge_inputs():
    # never raises; degenerate inputs take the reference's branches
    ColorExtractor.set_rng_seed(1)
    np.random.seed(1)
    one = ColorExtractor.extract_colors(np.full((8, 8, 3), 255, np.uint8), 5)
    assert one.metadata["success"] is True
    none = ColorExtractor.extract_colors(None, 5)          # 100x100 black default image
    assert none.metadata["success"] is True
    dev_mode = ColorExtractor.noise_mode
    try:
        ColorExtractor.noise_mode = "device"
        cf = ColorExtractor.extract_colors(design_image(90, 120, 5), 5)
        assert cf.metadata["success"] is True and cf.primary.startswith("#")
    finally:
        ColorExtractor.noise_mode = dev_mode
    centers, labels = ColorExtractor._get_dominant_colors(np.array([[1, 2, 3], [1, 2, 3]], np.uint8), 5)
    assert centers.tolist() == [[1, 2, 3]] and labels.tolist() == [0]


def test_preprocessing_and_transforms(golden):
    meta, arrays = golden
    big = design_image(600, 2400, 5)
    ok, png = cv2.imencode(".png", big)
    out = asyncio.run(validate_and_preprocess_image(png.tobytes(), "t", "auto"))
    r = [x for x in meta["resize"] if x["name"] == "auto_600x2400_s5"][0]
    assert list(out.shape) == r["out_shape"] and sha(out) == r["out_sha256"]
    perf = asyncio.run(validate_and_preprocess_image(png.tobytes(), "t", "performance"))     # utils.py:136-143
    assert np.array_equal(perf, cv2.resize(big, (1000, 250), interpolation=cv2.INTER_LINEAR))
    huge = design_image(1100, 4400, 9)
    ok2, png2 = cv2.imencode(".png", huge)
    hq = asyncio.run(validate_and_preprocess_image(png2.tobytes(), "t", "high_quality"))       # utils.py:128-135
    assert np.array_equal(hq, cv2.resize(huge, (4000, 1000), interpolation=cv2.INTER_LANCZOS4))
    same = asyncio.run(validate_and_preprocess_image(png.tobytes(), "t", "none"))
    assert np.array_equal(same, big)
    with pytest.raises(Exception) as e:
        asyncio.run(validate_and_preprocess_image(b"not an image", "t", "auto"))
    assert getattr(e.value, "status_code", None) == 400
    big2 = design_image(1300, 2200, 6)
    r2 = [x for x in meta["resize"] if x["name"] == "fit_1300x2200_s6"][0]
    out2 = ImageProcessor.resize_image(big2)
    assert list(out2.shape) == r2["out_shape"] and sha(out2) == r2["out_sha256"]
    assert np.array_equal(ImageProcessor.load_cv2_image(png.tobytes()), big)
    with pytest.raises(ValueError):
        ImageProcessor.load_cv2_image(b"")
    with pytest.raises(TypeError):
        ImageProcessor.resize_image("nope")
    src3 = design_image(512, 768, 7)
    assert np.array_equal(ImageTransformer.resize(src3.copy(), width=384, height=256), arrays["it_half_512x768_s7/out"])
    assert np.array_equal(ImageTransformer.generate_thumbnail(src3.copy(), size=(192, 128)), arrays["it_quarter_512x768_s7/out"])
    assert np.array_equal(ImageTransformer.resize(src3.copy(), width=500, height=300), arrays["it_frac_512x768_s7/out"])
    src4 = design_image(120, 200, 8)
    assert np.array_equal(ImageTransformer.adjust_brightness_contrast(src4.copy(), brightness=1.2, contrast=0.9),
                          arrays["transform/bc_1.2_0.9"])
    assert np.array_equal(ImageTransformer.apply_filter(src4.copy(), "gaussian_blur"), arrays["transform/gaussian_blur"])
    with pytest.raises(ValueError):
        ImageTransformer.apply_filter(src4, "nope")


def _resize_modes_golden():
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_resize_modes.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(_resize_modes_golden()["cases"]))
def test_preprocessing_modes_match_reference_golden(name):
    """validate_and_preprocess_image in the `performance` and `high_quality` modes on the GPU vs the outputs of
    the reference itself (tests/golden/golden_resize_modes.json)."""
    c = _resize_modes_golden()["cases"][name]
    img = design_image(c["h"], c["w"], c["seed"])
    assert sha(img) == c["input_sha256"]
    ok, png = cv2.imencode(".png", img)
    out = asyncio.run(validate_and_preprocess_image(png.tobytes(), "t", c["mode"]))
    assert list(out.shape) == c["out_shape"] and sha(out) == c["out_sha256"]
