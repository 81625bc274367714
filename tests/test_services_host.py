"""Host-side logic of the drop-in services (CPU only; no compute calls)."""
import numpy as np
import pytest
from PIL import Image

from oracle import cvops, load_reference
from low_level_feature_extraction_b200.services import ColorExtractor, ColorFeatures
from low_level_feature_extraction_b200.services.color_extractor import _CV_RNG_COEFF  # noqa: F401


def test_helpers():
    assert ColorExtractor.rgb_to_hex((255, 0, 16)) == "#ff0010"
    assert ColorExtractor.hex_to_rgb("#ff0010") == (255, 0, 16)
    assert ColorExtractor.is_light_color((255, 255, 255)) and not ColorExtractor.is_light_color((10, 10, 10))
    assert abs(ColorExtractor.get_contrast_ratio("#ffffff", "#000000") - 21.0) < 1e-9
    cf = ColorFeatures(primary="#112233", background="#FFFFFF", accent=["#000000"] * 3, metadata={"success": True})
    assert cf.primary == "#112233"
    with pytest.raises(Exception):
        ColorFeatures(primary="nothex", background="#FFFFFF", accent=[])


def test_rng_state_advance_matches_cv_rng():
    ColorExtractor.set_rng_seed(12345)
    r = cvops.CvRNG(12345)
    ColorExtractor._advance_rng(37)
    for _ in range(37):
        r.next()
    assert ColorExtractor._rng_state == r.state
    ColorExtractor.set_rng_seed(0)
    assert ColorExtractor._rng_state == 0xFFFFFFFF


INPUTS = {
    "bgr": lambda r: r.integers(0, 256, (12, 9, 3), dtype=np.uint8),
    "gray2d": lambda r: r.integers(0, 256, (12, 9), dtype=np.uint8),
    "gray1ch": lambda r: r.integers(0, 256, (12, 9, 1), dtype=np.uint8),
    "bgra": lambda r: r.integers(0, 256, (12, 9, 4), dtype=np.uint8),
    "chw_quirk": lambda r: r.integers(0, 256, (3, 20, 7), dtype=np.uint8),
    "float": lambda r: r.random((12, 9, 3)).astype(np.float32),
    "flat_square": lambda r: r.integers(0, 256, (6 * 6 * 3,), dtype=np.uint8),
    "flat_bad": lambda r: r.integers(0, 256, (50,), dtype=np.uint8),
    "empty": lambda r: np.zeros((0, 3), np.uint8),
    "five_ch": lambda r: r.integers(0, 256, (12, 9, 5), dtype=np.uint8),
    "pil_rgb": lambda r: Image.fromarray(r.integers(0, 256, (12, 9, 3), dtype=np.uint8), "RGB"),
    "pil_rgba": lambda r: Image.fromarray(r.integers(0, 256, (12, 9, 4), dtype=np.uint8), "RGBA"),
    "pil_l": lambda r: Image.fromarray(r.integers(0, 256, (12, 9), dtype=np.uint8), "L"),
    "none": lambda r: None,
}


@pytest.mark.skipif(not load_reference.available(), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("kind", sorted(INPUTS))
def test_process_image_matches_live_reference(kind):
    ref = load_reference.load()["ColorExtractor"]
    x = INPUTS[kind](np.random.default_rng(1))
    a = ColorExtractor._process_image(x if not isinstance(x, np.ndarray) else x.copy())
    b = ref._process_image(x if not isinstance(x, np.ndarray) else x.copy())
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), kind


def test_process_image_without_reference():
    r = np.random.default_rng(2)
    bgr = INPUTS["bgr"](r)
    assert np.array_equal(ColorExtractor._process_image(bgr), bgr[..., ::-1])
    assert ColorExtractor._process_image(None).shape == (100, 100, 3)
    g = INPUTS["gray2d"](r)
    assert np.array_equal(ColorExtractor._process_image(g), np.stack([g, g, g], -1))


BAD_INPUTS = {
    "gray2d": lambda r: r.integers(0, 256, (40, 120), dtype=np.uint8),
    "two_ch": lambda r: r.integers(0, 256, (40, 120, 2), dtype=np.uint8),
    "five_ch": lambda r: r.integers(0, 256, (40, 120, 5), dtype=np.uint8),
    "float": lambda r: r.random((40, 120, 3)).astype(np.float32),
    "u16": lambda r: r.integers(0, 60000, (40, 120, 3), dtype=np.uint16),
    "empty": lambda r: np.zeros((0, 0, 3), np.uint8),
    "list": lambda r: [[1, 2, 3]],
    "none": lambda r: None,
}


@pytest.mark.parametrize("kind", sorted(BAD_INPUTS))
def test_service_input_contract_raises_like_cv2(kind):
    """Inputs the reference's cv2 calls reject raise cv2.error in the drop-in too (before any GPU work), checked
    against the live reference classes where they are available."""
    import cv2
    from low_level_feature_extraction_b200.services import FontDetector, ShadowAnalyzer, ShapeAnalyzer

    x = BAD_INPUTS[kind](np.random.default_rng(3))
    for fn in (ShapeAnalyzer.preprocess_image, ShadowAnalyzer.analyze_shadow_level, FontDetector.preprocess_image):
        with pytest.raises(cv2.error):
            fn(x)
    if load_reference.available():
        ref = load_reference.load()
        for fn in (ref["ShapeAnalyzer"].preprocess_image, ref["ShadowAnalyzer"].analyze_shadow_level):
            with pytest.raises(cv2.error):
                fn(x)


def test_batched_palette_tail_equals_per_image_tail():
    """`_palettes_from_batch` == `_palette_from_clusters` image by image, including tied cluster sizes (argsort's tie
    order), pure white / black centres, palettes that lose every colour, and images with fewer colours than clusters."""
    r = np.random.default_rng(7)
    n, k = 600, 5
    centers = r.integers(0, 256, (n, k, 3)).astype(np.float32) + r.random((n, k, 3)).astype(np.float32) * 0.99
    sizes = r.integers(1, 6, (n, k)).astype(np.int32)             # many ties
    k_used = np.full(n, k, np.int32)
    centers[::7, 0] = 255.2                                        # white
    centers[::11, 1] = 0.4                                         # black
    centers[5] = np.array([[255.0] * 3, [0.0] * 3, [255.9] * 3, [0.9] * 3, [255.0] * 3])   # nothing left
    centers[::13, 2] = centers[::13, 3]                            # duplicate colours
    k_used[3], k_used[4] = 2, 1
    sizes[3, 2:] = 0
    sizes[4, 1:] = 0
    got = ColorExtractor._palettes_from_batch(centers, k_used, sizes)
    for i in range(n):
        kk = int(k_used[i])
        want = ColorExtractor._palette_from_clusters(centers[i, :kk].astype(np.uint8),
                                                     sizes[i, :kk].astype(np.int64) if kk > 1 else None)
        assert (got[i].primary, got[i].background, list(got[i].accent), got[i].metadata) == \
               (want.primary, want.background, list(want.accent), want.metadata), i
