"""oracle/cvops.py vs the OpenCV binary the reference calls (CPU only).

cv2 is part of the image (also on the GPU box), so this pins the numpy
restatement on inputs the golden fixtures do not cover: odd sizes, tiny
images, adversarial noise, every INTER_AREA code path, empty clusters."""
import numpy as np
import cv2
import pytest

from oracle import cvops, refpath
from low_level_feature_extraction_b200.synth import design_image, noise_image

SHAPES = [(64, 96), (53, 37), (11, 11), (7, 5), (5, 64), (1, 40), (40, 1), (3, 3), (2, 2), (135, 257)]


@pytest.mark.parametrize("shape", SHAPES)
def test_gray_blur(shape):
    img = noise_image(*shape, seed=shape[0])
    g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(cvops.bgr2gray(img), g)
    assert np.array_equal(cvops.gaussian_blur5(g), cv2.GaussianBlur(g, (5, 5), 0))
    assert np.array_equal(cvops.gaussian_blur5(img), cv2.GaussianBlur(img, (5, 5), 0))


def test_gray_all_colors_sampled():
    r = np.random.default_rng(0)
    img = r.integers(0, 256, (1, 1 << 18, 3), dtype=np.uint8)
    assert np.array_equal(cvops.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("shape", [s for s in SHAPES if min(s) >= 2])
def test_canny_dilate(shape):
    r = np.random.default_rng(shape[1])
    g = cv2.GaussianBlur(r.integers(0, 256, shape, dtype=np.uint8), (5, 5), 0)
    for src in (g, r.integers(0, 256, shape, dtype=np.uint8)):
        e = cv2.Canny(src, 50, 150)
        assert np.array_equal(cvops.canny(src, 50, 150), e)
        assert np.array_equal(cvops.dilate3(e), cv2.dilate(e, np.ones((3, 3), np.uint8), iterations=1))
        w, s = cvops.canny_nms(src, 50, 150)
        assert np.array_equal(cvops.hysteresis(w, s), cvops.hysteresis_iterative(w, s))


def test_shape_mask_design():
    img = design_image(240, 320, 9)
    assert np.array_equal(cvops.shape_mask(img), refpath.shape_mask(img))


@pytest.mark.parametrize("shape", [(64, 96), (53, 40), (11, 16), (135, 256), (30, 1920), (45, 77), (33, 9), (20, 203),
                                   (17, 1001), (12, 1366), (9, 1444), (7, 1921), (40, 67), (40, 69), (2, 13)]
                         + [(6, w) for w in range(2, 41)])
def test_adaptive(shape):
    # any width: the restatement follows OpenCV's vector body and its tail columns, float for float
    r = np.random.default_rng(shape[0])
    g = cv2.GaussianBlur(r.integers(0, 256, shape, dtype=np.uint8), (5, 5), 0)
    cvf = cv2.GaussianBlur(g.astype(np.float32), (11, 11), 0, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(cvops.gauss11_f32(g), cvf)
    out = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, 2)
    assert np.array_equal(cvops.adaptive_threshold(g, 2), out)


def test_gaussian_kernel():
    assert np.array_equal(cvops.gaussian_kernel_f32(11), cv2.getGaussianKernel(11, 0).astype(np.float32).ravel())


@pytest.mark.parametrize("seed", range(6))
def test_otsu(seed):
    r = np.random.default_rng(seed)
    if seed == 0:
        g = np.full((40, 120), 77, np.uint8)          # degenerate constant image -> t = 0
    elif seed == 1:
        g = cvops.bgr2gray(design_image(120, 200, seed))
    else:
        g = np.clip(r.normal(r.integers(60, 200), 40, (90, 130)), 0, 255).astype(np.uint8)
    t, b = cv2.threshold(g, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    t2, b2 = cvops.otsu_binary(g)
    assert t2 == int(t) and np.array_equal(b, b2)
    img = np.stack([g, g, g], -1)
    assert np.array_equal(cvops.text_mask(img), refpath.text_mask(img))


def test_otsu_tiny_class_uses_flt_epsilon():
    # OpenCV skips classes lighter than FLT_EPSILON: 1 px of 16.7M is ignored, 3 px are not
    g = np.full((4096, 4096), 200, np.uint8)
    g[0, 0] = 50
    assert cvops.otsu_threshold(g) == int(cv2.threshold(g, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0]) == 0
    g[0, 0:3] = 50
    assert cvops.otsu_threshold(g) == int(cv2.threshold(g, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0]) == 50


@pytest.mark.parametrize("case", [((216, 384, 3), (200, 112)), ((120, 240, 3), (200, 100)), ((150, 210), (200, 142)),
                                  ((128, 128, 3), (64, 64)), ((128, 256, 3), (64, 32)), ((96, 96), (12, 12)),
                                  ((90, 120, 3), (40, 30)), ((100, 300, 3), (299, 99)), ((64, 64, 3), (64, 64))])
def test_resize_area(case):
    shp, (dw, dh) = case
    src = np.random.default_rng(dw).integers(0, 256, shp, dtype=np.uint8)
    assert np.array_equal(cvops.resize_area(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA))


LINEAR_CASES = [((1080, 1920, 3), (1000, 562)), ((300, 500, 3), (250, 150)), ((1200, 1000), (833, 1000)),
                ((77, 131, 3), (60, 30)), ((101, 203, 3), (100, 50)), ((64, 64, 3), (32, 32)), ((40, 60, 3), (90, 70)),
                ((33, 47), (47, 33)), ((50, 50, 3), (50, 50))]


@pytest.mark.parametrize("case", LINEAR_CASES)
def test_resize_linear(case):
    """cv2's fixed-point INTER_LINEAR (the `performance` preprocessing mode, utils.py:136-143), incl. the
    exact-2x case that OpenCV reroutes to INTER_AREA, up-scaling and the identity."""
    shp, (dw, dh) = case
    src = np.random.default_rng(dw + dh).integers(0, 256, shp, dtype=np.uint8)
    assert np.array_equal(cvops.resize_linear(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("case", [((300, 500, 3), (250, 150)), ((1200, 1000), (833, 1000)), ((77, 131, 3), (60, 30)),
                                  ((64, 64, 3), (32, 32)), ((40, 60, 3), (90, 70)), ((33, 47), (47, 33)), ((50, 50, 3), (50, 50)),
                                  ((9, 7, 3), (5, 4))])
def test_resize_lanczos4(case):
    """cv2's fixed-point INTER_LANCZOS4 (the `high_quality` preprocessing mode, utils.py:128-135)."""
    shp, (dw, dh) = case
    src = np.random.default_rng(dw * 3 + dh).integers(0, 256, shp, dtype=np.uint8)
    assert np.array_equal(cvops.resize_lanczos4(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4))


def test_performance_resize_shape_matches_reference_rule():
    assert cvops.performance_resize_shape(1080, 1920) == (1000, 562)
    assert cvops.performance_resize_shape(2160, 3840) == (1000, 562)
    assert cvops.performance_resize_shape(800, 1000) is None


@pytest.mark.parametrize("alpha", [1.2, 0.9, 1.5, 0.5, 1.3, 1.0, 2.5])
def test_convert_scale_abs(alpha):
    x = np.arange(256, dtype=np.uint8).reshape(1, -1)
    assert np.array_equal(cvops.convert_scale_abs(x, alpha), cv2.convertScaleAbs(x, alpha=alpha, beta=0))


def test_unique_is_np_unique():
    px = np.random.default_rng(0).integers(0, 40, (5000, 3), dtype=np.uint8)
    assert np.array_equal(cvops.unique_colors(px), np.unique(px, axis=0))
    u, c = cvops.unique_colors_counts(px)
    u2, c2 = np.unique(px, axis=0, return_counts=True)
    assert np.array_equal(u, u2) and np.array_equal(c, c2)


@pytest.mark.parametrize("seed,k", [(1, 5), (7, 16), (12345, 5)])
def test_kmeans_pp_and_full(seed, k):
    img = design_image(96, 128, seed % 5)
    data = np.float32(np.unique(img.reshape(-1, 3), axis=0))
    # kmeans++ init: sequential walk == prefix formulation, and == cv2 (labels after 1 iteration)
    a, ia = cvops.centers_pp(data, k, cvops.CvRNG(seed), sequential=True)
    b, ib = cvops.centers_pp(data, k, cvops.CvRNG(seed), sequential=False)
    assert ia == ib
    cv2.setRNGSeed(seed)
    _, labels, _ = cv2.kmeans(data, k, None, (cv2.TERM_CRITERIA_MAX_ITER, 1, 0), 1, cv2.KMEANS_PP_CENTERS)
    assert np.array_equal(cvops.assign(data, a)[0], labels.ravel())
    # full 10-attempt kmeans
    cv2.setRNGSeed(seed)
    comp, labels, centers = cv2.kmeans(data, k, None, refpath.KMEANS_CRITERIA, 10, cv2.KMEANS_PP_CENTERS)
    c2, l2, ce2 = cvops.cv_kmeans(data, k, cvops.CvRNG(seed))
    assert np.array_equal(l2, labels.ravel()) and np.array_equal(ce2, centers)
    assert abs(c2 - comp) <= 1e-9 * max(1.0, comp)


@pytest.mark.parametrize("seed", range(4))
def test_lloyd_seeded_and_empty_cluster_repair(seed):
    r = np.random.default_rng(seed)
    data = np.float32(np.unique(r.integers(0, 256, (3000, 3), dtype=np.uint8), axis=0))
    k = 6
    init = data[r.choice(len(data), k, replace=False)].copy()
    init[4] = init[1]      # duplicate centres -> cluster 4 starts empty (strict '<' keeps the lower index)
    init[5] = init[1]
    c_cv, l_cv, _ = refpath.kmeans_pixels(data, init)
    c, l, it, comp = cvops.lloyd_cv(data, init)
    assert np.array_equal(l, l_cv) and np.array_equal(c, c_cv)


def test_lloyd_exact_matches_cv_when_sums_are_small():
    # exact integer sums == cv2's sequential f32 sums while every channel sum < 2^24
    img = design_image(96, 128, 3)
    u8 = np.unique(img.reshape(-1, 3), axis=0)
    r = np.random.default_rng(5)
    init = np.float32(u8[r.choice(len(u8), 5, replace=False)])
    c_cv, l_cv, _ = refpath.kmeans_pixels(np.float32(u8), init)
    c, l, it, sums, cnt = cvops.lloyd_exact(u8, init)
    assert np.array_equal(l, l_cv)
    assert np.allclose(c, c_cv, rtol=1e-6, atol=1e-4)
    # weighted (unique colours + counts) formulation == raw pixel list
    px = img.reshape(-1, 3)
    uu, cc = cvops.unique_colors_counts(px)
    c1, l1, it1, s1, n1 = cvops.lloyd_exact(px, init)
    c2, l2, it2, s2, n2 = cvops.lloyd_exact(uu, init, weights=cc)
    assert it1 == it2 and np.array_equal(c1, c2) and np.array_equal(s1, s2) and np.array_equal(n1, n2)


def test_extract_colors_matches_cv2_port():
    img = design_image(120, 160, 2)
    for seed in (3, 11):
        np.random.seed(seed)
        cv2.setRNGSeed(seed)
        ref = refpath.extract_colors(img, 5)
        noise = cvops.make_noise((img.shape[0] * img.shape[1], 3), seed)
        assert cvops.extract_colors(img, 5, noise, seed) == ref


def test_kmeans_max_count_is_clamped_to_100():
    """cv::kmeans clamps criteria.maxCount to [2, 100]: the reference's 200 (color_extractor.py:190) means 100.
    A slow-converging list with eps = 0: cv2 returns the same result for 100, 101 and 200, another one for 99,
    and the restatement follows."""
    r = np.random.default_rng(5)
    data = r.integers(0, 256, (20000, 3)).astype(np.float32)
    res = {}
    for mc in (99, 100, 200):
        cv2.setRNGSeed(7)
        res[mc] = cv2.kmeans(data, 16, None, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, mc, 0.0), 1,
                             cv2.KMEANS_PP_CENTERS)
    assert res[100][0] == res[200][0] and np.array_equal(res[100][1], res[200][1])
    assert res[99][0] != res[100][0]
    comp, labels, centers = cvops.cv_kmeans(data, 16, cvops.CvRNG(7), attempts=1, max_iter=200, eps=0.0)
    assert np.array_equal(labels, res[200][1].ravel()) and np.array_equal(centers, res[200][2])
    assert abs(comp - res[200][0]) <= 1e-9 * res[200][0]
