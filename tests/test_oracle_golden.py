"""oracle/cvops.py (numpy restatement) vs the outputs of the UNMODIFIED reference
captured in tests/golden/ (CPU only)."""
import hashlib

import numpy as np
import pytest

from oracle import cvops

CASES = ["design_270x480_s1", "design_360x640_s2", "noise_96x160_s3", "design_101x203_s4"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_shape_mask(golden, golden_inputs, name):
    _, arrays = golden
    assert np.array_equal(cvops.shape_mask(golden_inputs[name]), arrays[name + "/shape_mask"])


@pytest.mark.parametrize("name", CASES)
def test_shadow(golden, golden_inputs, name):
    meta, arrays = golden
    blurred, mask, s, n, level = cvops.shadow_parts(golden_inputs[name])
    assert np.array_equal(blurred, arrays[name + "/shadow_blurred"])
    assert np.array_equal(mask, arrays[name + "/shadow_thresh"])
    assert level == meta["cases"][name]["shadow_level"]


@pytest.mark.parametrize("name", CASES)
def test_text_and_font_masks(golden, golden_inputs, name):
    _, arrays = golden
    img = golden_inputs[name]
    if img.shape[0] >= 30 and img.shape[1] >= 100:
        assert np.array_equal(cvops.text_mask(img), arrays[name + "/text_mask"])
    assert np.array_equal(cvops.font_mask(img), arrays[name + "/font_mask"])


def test_colors(golden, golden_inputs):
    meta, arrays = golden
    for c in meta["colors"]:
        img = golden_inputs[c["case"]]
        noise = cvops.make_noise((img.shape[0] * img.shape[1], 3), c["seed"])
        px = cvops.apply_noise(cvops.bgr2rgb(img).reshape(-1, 3), noise)
        assert len(cvops.unique_colors(px)) == c["n_unique"]
        centers, labels = cvops.dominant_colors(px, c["k"], cvops.CvRNG(c["seed"]))
        assert np.array_equal(centers, arrays[c["tag"] + "/centers"]), c["tag"]
        assert np.array_equal(labels, arrays[c["tag"] + "/labels"]), c["tag"]
        out = cvops.extract_colors(img, c["k"], noise, c["seed"])
        for key in ("primary", "background", "accent"):
            assert out[key] == c["result"][key], (c["tag"], key)


def test_resize(golden):
    from low_level_feature_extraction_b200.synth import design_image

    meta, arrays = golden
    srcs = {"auto_600x2400_s5": design_image(600, 2400, 5), "fit_1300x2200_s6": design_image(1300, 2200, 6)}
    src3 = design_image(512, 768, 7)
    for r in meta["resize"]:
        src = srcs.get(r["name"], src3)
        assert sha(src) == r["in_sha256"]
        dh, dw = r["out_shape"][:2]
        if r["name"].startswith("auto"):
            assert cvops.auto_resize_shape(*src.shape[:2]) == (dh, dw)
        if r["name"].startswith("fit"):
            assert cvops.fit_resize_shape(*src.shape[:2]) == (dh, dw)
        out = cvops.resize_area(src, dw, dh)
        assert sha(out) == r["out_sha256"], r["name"]


def test_transform(golden):
    from low_level_feature_extraction_b200.synth import design_image

    meta, arrays = golden
    src = design_image(120, 200, 8)
    assert sha(src) == meta["transform"]["input_sha256"]
    assert np.array_equal(cvops.adjust_brightness_contrast(src, 1.2, 0.9), arrays["transform/bc_1.2_0.9"])
    assert np.array_equal(cvops.adjust_brightness_contrast(src, 1.3, 1.5), arrays["transform/bc_1.3_1.5"])
    assert np.array_equal(cvops.gaussian_blur5(src), arrays["transform/gaussian_blur"])


def _resize_modes_golden():
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_resize_modes.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(_resize_modes_golden()["cases"]))
def test_resize_modes_oracle_matches_reference(name):
    """oracle restatements of the `performance` / `high_quality` modes vs outputs of the reference itself
    (oracle/make_golden_resize_modes.py)."""
    from low_level_feature_extraction_b200.synth import design_image

    c = _resize_modes_golden()["cases"][name]
    img = design_image(c["h"], c["w"], c["seed"])
    assert sha(img) == c["input_sha256"]
    h, w = img.shape[:2]
    if c["mode"] == "performance":
        shp = cvops.performance_resize_shape(h, w)
        out = img if shp is None else cvops.resize_linear(img, shp[0], shp[1])
    else:
        max_dim = 4000
        if max(h, w) > max_dim:
            s = max_dim / max(h, w)
            out = cvops.resize_lanczos4(img, int(w * s), int(h * s))
        else:
            out = img
    assert list(out.shape) == c["out_shape"] and sha(out) == c["out_sha256"]
