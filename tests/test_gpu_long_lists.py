"""cv2.kmeans on colour lists long enough that OpenCV's sequential float32 centre sums round (a cluster's channel
sum >= 2^24, i.e. more than ~65 793 members): the kernel reproduces those sums operation for operation
(k_kmeans_fast.cu:seq_f32_sums3), so labels and centres stay bit-identical to cv2.kmeans -- including SURVEY 8(d)'s
adversarial frame (uniform noise at 1080p, U ~ 1.95 M, K = 5)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import cv2  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    import low_level_feature_extraction_b200 as pkg
    return pkg.engine(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_both(eng, keys_u32, k, seed, attempts=10):
    uniq = np.stack([keys_u32 >> 16, (keys_u32 >> 8) & 255, keys_u32 & 255], 1).astype(np.float32)
    cv2.setRNGSeed(seed)
    comp_cv, lab_cv, cen_cv = cv2.kmeans(uniq, k, None, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 200, 0.2),
                                         attempts, cv2.KMEANS_PP_CENTERS)
    n = len(keys_u32)
    centers, labels, comp, kused = eng.kmeans_unique(dev(keys_u32.astype(np.int64).astype(np.int32)[None]),
                                                     dev(np.array([n], np.int32)), k, seed, attempts=attempts)
    return (comp_cv, lab_cv.ravel(), cen_cv), (float(comp[0]), labels[0, :n].cpu().numpy(), centers[0].cpu().numpy(),
                                              int(eng.last_status[0]))


@pytest.mark.parametrize("case", [(100_000, 2, 3, False), (300_000, 2, 11, True), (500_000, 3, 5, True), (700_000, 8, 7, True)])
def test_long_lists_match_cv2(eng, case):
    n, k, seed, expect_long = case
    r = np.random.default_rng(seed)
    keys = np.sort(r.choice(1 << 24, n, replace=False)).astype(np.uint32)
    (comp_cv, lab_cv, cen_cv), (comp, lab, cen, status) = run_both(eng, keys, k, seed, attempts=3)
    assert np.array_equal(cen, cen_cv), (cen, cen_cv)
    assert np.array_equal(lab, lab_cv)
    assert abs(comp - comp_cv) <= 1e-9 * comp_cv
    assert bool(status & 1) == expect_long


def test_skewed_list_one_long_channel(eng):
    """One cluster's R sum is long while its B sum is not, and a dark cluster stays exact: mixed exact / sequential sums."""
    r = np.random.default_rng(1)
    a = np.stack([r.integers(200, 256, 150_000), r.integers(0, 256, 150_000), r.integers(0, 40, 150_000)], 1)
    b = np.stack([r.integers(0, 30, 40_000), r.integers(0, 256, 40_000), r.integers(0, 256, 40_000)], 1)
    px = np.unique(np.concatenate([a, b]).astype(np.uint8), axis=0).astype(np.uint32)
    keys = (px[:, 0] << 16) | (px[:, 1] << 8) | px[:, 2]
    (comp_cv, lab_cv, cen_cv), (comp, lab, cen, status) = run_both(eng, keys, 2, 9, attempts=2)
    assert np.array_equal(cen, cen_cv) and np.array_equal(lab, lab_cv) and status & 1


def test_adversarial_noise_frame_matches_cv2(eng):
    """rng.integers(0, 256, (1080, 1920, 3)), U ~ 1.95 M unique colours, K = 5, the reference's call
    (10 kmeans++ attempts) -- through the service method a user calls."""
    from low_level_feature_extraction_b200.services import ColorExtractor

    img = np.random.default_rng(2024).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    px = img.reshape(-1, 3)
    uniq = np.unique(px, axis=0)
    assert len(uniq) > 1_900_000
    cv2.setRNGSeed(77)
    _, lab_cv, cen_cv = cv2.kmeans(np.float32(uniq), 5, None, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 200, 0.2),
                                   10, cv2.KMEANS_PP_CENTERS)
    ColorExtractor.set_rng_seed(77)
    centers, labels = ColorExtractor._get_dominant_colors(px, 5)      # color_extractor.py:173-201 on the GPU
    assert ColorExtractor.last_status & 1
    assert np.array_equal(centers, cen_cv.astype(np.uint8))
    assert np.array_equal(labels, lab_cv.ravel())
