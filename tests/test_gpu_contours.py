"""Device contour finder (llfe_contours_external*, csrc/k_contours.cu) against cv2.findContours(mask, RETR_EXTERNAL,
CHAIN_APPROX_SIMPLE) itself: same contours, same points, same order; through the C ABI."""
import cv2
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import refpath  # noqa: E402
from low_level_feature_extraction_b200 import contours as ct  # noqa: E402
from low_level_feature_extraction_b200.services import FontDetector, ShapeAnalyzer, _runtime  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image, noise_image  # noqa: E402
from test_contours_oracle import random_mask  # noqa: E402


def cv_all(m):
    ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return list(ref)


def check(mask, min_area2=0):
    ctx = _runtime.context()
    headers, points = ct.find_external_host(ctx, mask, min_area2)
    ref = cv_all(mask)
    hs = ct.in_cv2_order(headers)
    assert len(hs) == len(ref)
    for h, c in zip(hs, ref):
        assert (h["start"] % mask.shape[1], h["start"] // mask.shape[1]) == tuple(c[0][0])
        assert h["npts"] == len(c)
        assert abs(int(h["area2"])) == int(round(2 * cv2.contourArea(c)))
        x, y, w, hh = cv2.boundingRect(c)
        assert (h["minx"], h["miny"], h["maxx"] - h["minx"] + 1, h["maxy"] - h["miny"] + 1) == (x, y, w, hh)
        if abs(int(h["area2"])) >= min_area2:
            assert h["offset"] >= 0
            assert np.array_equal(points[h["offset"]:h["offset"] + h["npts"]], c.reshape(-1, 2))
        else:
            assert h["offset"] == -1
    got = ct.to_cv2_contours(headers, points)
    want = [c for c in ref if int(round(2 * cv2.contourArea(c))) >= min_area2]
    assert len(got) == len(want) and all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(got, want))


def test_random_small_masks():
    rng = np.random.default_rng(10)
    for _ in range(300):
        h, w = rng.integers(1, 80, 2)
        check(random_mask(rng, int(h), int(w)))


@pytest.mark.parametrize("shape", [(1, 1), (1, 77), (77, 1), (31, 32), (32, 33), (33, 31), (64, 64), (65, 127), (200, 333)])
def test_shapes_and_edge_cases(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    check(np.zeros(shape, np.uint8))
    check(np.full(shape, 255, np.uint8))
    check((rng.random(shape) < 0.5).astype(np.uint8))          # any non-zero byte is foreground
    m = np.zeros(shape, np.uint8)
    m[0, :] = m[-1, :] = 255
    m[:, 0] = m[:, -1] = 255                                    # a ring on the frame: everything inside is internal
    if h > 6 and w > 6:
        m[3:-3, 3:-3] = (rng.random((h - 6, w - 6)) < 0.4) * 255
    check(m)
    check(255 - m)


def test_nested_rings_and_spiral():
    m = np.zeros((120, 150), np.uint8)
    for k in range(0, 50, 6):
        cv2.rectangle(m, (10 + k, 10 + k), (139 - k, 109 - k), 255, 1)
    check(m)
    s = np.zeros((101, 101), np.uint8)                          # a one-pixel spiral wall: one long border
    x0, y0, x1, y1 = 0, 0, 100, 100
    while x1 - x0 > 4:
        s[y0, x0:x1 + 1] = 255
        s[y0:y1 + 1, x1] = 255
        s[y1, x0 + 2:x1 + 1] = 255
        s[y0 + 2:y1 + 1, x0 + 2] = 255
        x0, y0, x1, y1 = x0 + 2, y0 + 2, x1 - 2, y1 - 2
    check(s)
    check(255 - s)


def test_diagonals_and_circles():
    m = np.zeros((300, 400), np.uint8)
    cv2.circle(m, (100, 100), 80, 255, 2)
    cv2.circle(m, (100, 100), 30, 255, -1)
    cv2.ellipse(m, (280, 150), (100, 60), 30, 0, 360, 255, 3)
    cv2.line(m, (5, 290), (390, 200), 255, 1)
    cv2.line(m, (5, 200), (390, 295), 255, 2)
    check(m)
    check(m, 200)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_design_masks_1080p(seed):
    img = design_image(1080, 1920, seed)
    mask = refpath.shape_mask(img)
    check(mask)
    check(mask, ct.REFERENCE_MIN_AREA2)
    check(refpath.font_mask(img))                               # dense adaptive-threshold mask, thousands of contours


def test_noise_mask_large():
    rng = np.random.default_rng(5)
    for dens in (0.2, 0.45, 0.8):
        check((rng.random((540, 700)) < dens).astype(np.uint8) * 255)
    check(refpath.shape_mask(noise_image(360, 640, 3)))


def test_batched_device_call(llfe):
    eng = llfe.engine(0)
    masks = np.stack([refpath.shape_mask(design_image(270, 480, s)) for s in range(5)])
    d = torch.from_numpy(masks).cuda()
    mc = 2048
    hdr, pts, cnt = eng.contours_external(d, 200, mc, 1 << 15)
    for i in range(5):
        c = cnt[i].cpu().numpy()
        assert c[0] <= mc and not c[2]
        headers = hdr[i, :c[0]].cpu().numpy().view(ct.HEADER).reshape(-1)
        got = ct.to_cv2_contours(headers, pts[i].cpu().numpy())
        want = [x for x in cv_all(masks[i]) if cv2.contourArea(x) >= 100]
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))


def test_capacity_retry():
    rng = np.random.default_rng(3)
    m = (rng.random((300, 300)) < 0.2).astype(np.uint8) * 255      # ~10 k contours > the first capacity guess
    assert len(cv_all(m)) > 4096
    check(m)


def test_services_use_device_contours(monkeypatch):
    img = design_image(360, 640, 2)

    def boom(*a, **k):
        raise AssertionError("cv2.findContours must not be called by the drop-in services")

    want = refpath.analyze_shapes_from_mask(refpath.shape_mask(img), img.shape)
    regions_want = []
    for c in cv_all(refpath.font_mask(img)):                     # font_detector.py:57-66
        x, y, w, h = cv2.boundingRect(c)
        if 0.1 < w / float(h) < 15 and h > 8:
            regions_want.append((x, y, w, h))
    monkeypatch.setattr(cv2, "findContours", boom)
    assert ShapeAnalyzer.analyze_shapes(img) == want
    assert ShapeAnalyzer.shapes_from_mask(refpath.shape_mask(img), img.shape[1], img.shape[0]) == want
    assert FontDetector.detect_text_regions(FontDetector.preprocess_image(img)) == regions_want
    assert len(regions_want) > 0


def test_batch_analyzer_contours_and_overflow_fallback():
    from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig

    imgs = np.stack([design_image(270, 480, s) for s in range(5)] + [noise_image(270, 480, 9)])
    pinned = torch.from_numpy(imgs).pin_memory()
    an = BatchAnalyzer(0, 270, 480, BatchConfig(colors=False, shadows=False, contours=True, max_contours=2, max_points=4096,
                                                host_chunk=4))
    out = an.run_host(pinned)
    seen = set()
    for i in range(len(imgs)):
        mask = out["shape_mask"][i].numpy()
        assert np.array_equal(mask, refpath.shape_mask(imgs[i]))
        want = [c for c in cv_all(mask) if cv2.contourArea(c) >= 100]
        got = an.contours(out, i)
        seen.add(len(cv_all(mask)) > 2)
        if len(cv_all(mask)) > 2:
            assert got is None                                   # the caller redoes this image alone
        else:
            assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    assert seen == {True, False}


# ---- segments: long borders followed in parallel (k_ct_segments + the hopping leader) ---------------------------------
@pytest.fixture
def cut_shift():
    ctx = _runtime.context()

    def set_(shift, on=1):
        ctx.set_option("contour_cut_shift", shift)
        ctx.set_option("contour_segments", on)

    yield set_
    ctx.set_option("contour_cut_shift", 6)
    ctx.set_option("contour_segments", 1)


@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_segments_every_cut_spacing_on_small_masks(cut_shift, shift):
    """cut rows every 1 / 2 / 4 / 8 rows: nearly every step of a border is a hop, heads on the start row included"""
    cut_shift(shift)
    rng = np.random.default_rng(20 + shift)
    for k in range(150):
        h, w = rng.integers(1, 90, 2)
        m = random_mask(rng, int(h), int(w))
        if k % 4 == 0:
            m = cv2.dilate(m, np.ones((3, 3), np.uint8))
        check(m)
    for shape in [(1, 1), (1, 40), (40, 1), (33, 65)]:
        check(np.full(shape, 255, np.uint8))
        m = np.zeros(shape, np.uint8)
        m[::2, ::2] = 255
        check(m)


def test_segments_long_borders_and_capacity(cut_shift):
    """spirals / combs with borders of tens of thousands of pixels; segments longer than the recorded maximum (a long
    horizontal edge between two cut rows) are walked by the leader; same result with the segments switched off"""
    yy, xx = np.mgrid[0:700, 0:900]
    spiral = (((np.hypot(yy - 350, xx - 450) + 12 * np.arctan2(yy - 350, xx - 450) / np.pi) % 24) < 9).astype(np.uint8) * 255
    comb = np.zeros((600, 1500), np.uint8)
    comb[5:8, 3:-3] = 255
    comb[5:590, 3:-3:6] = 255
    zig = np.zeros((300, 1400), np.uint8)
    for x in range(2, 1390):
        zig[40 + (x * 7) % 200, x] = zig[41 + (x * 7) % 200, x] = 255
    zig = cv2.dilate(zig, np.ones((3, 3), np.uint8))
    masks = [spiral, comb, zig, refpath.shape_mask(design_image(1080, 1920, 1)), refpath.shape_mask(noise_image(270, 480, 3))]
    ctx = _runtime.context()
    for m in masks:
        for shift in (6, 4, 3):
            cut_shift(shift)
            check(m)
            check(m, 200)
        cut_shift(6)
        h1, p1 = ct.find_external_host(ctx, m, 200)
        cut_shift(6, on=0)
        h0, p0 = ct.find_external_host(ctx, m, 200)
        a, b = ct.to_cv2_contours(h1, p1), ct.to_cv2_contours(h0, p0)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
