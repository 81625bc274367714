"""CUDA kernels (through the C ABI) vs the CPU oracle, cv2 and the reference's golden outputs.

Bit-exact bar for every op here (all integer / byte work; the adaptive threshold's float32
intermediate is reproduced in OpenCV's exact operation order, so its mask is bit-exact too
for widths that are multiples of 8 -- see SURVEY.md A.5; other widths: <= 1 LSB on the float
mean, checked as mask mismatches only where |src - mean + C| could flip)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cvops, refpath  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image, noise_image  # noqa: E402

SHAPES = [(64, 96), (53, 37), (11, 11), (7, 5), (5, 64), (1, 40), (40, 1), (3, 3), (2, 2), (135, 257), (270, 480),
          (33, 2049)]


@pytest.fixture(scope="module")
def eng():
    import low_level_feature_extraction_b200 as pkg
    return pkg.engine(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("shape", SHAPES)
def test_gray_rgb_blur(eng, shape):
    img = noise_image(*shape, seed=shape[0] + shape[1])
    g = cvops.bgr2gray(img)
    assert np.array_equal(host(eng.bgr2gray(dev(img))), g)
    assert np.array_equal(host(eng.bgr2rgb(dev(img))), img[..., ::-1])
    assert np.array_equal(host(eng.gaussian_blur5(dev(g))), cvops.gaussian_blur5(g))
    assert np.array_equal(host(eng.gaussian_blur5(dev(img))), cvops.gaussian_blur5(img))
    assert np.array_equal(host(eng.gray_blur5(dev(img))), cvops.gaussian_blur5(g))


def test_gray_all_2pow24_colors(eng):
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([v & 255, (v >> 8) & 255, v >> 16], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(host(eng.bgr2gray(dev(img))), cvops.bgr2gray(img))


def test_batched_equals_single(eng):
    batch = np.stack([design_image(96, 160, s) for s in range(5)])
    out = host(eng.shape_mask(dev(batch)))
    for i in range(5):
        assert np.array_equal(out[i], cvops.shape_mask(batch[i]))


@pytest.mark.parametrize("shape", [s for s in SHAPES if min(s) >= 2])
def test_canny_dilate(eng, shape):
    r = np.random.default_rng(shape[1])
    noise = r.integers(0, 256, shape, dtype=np.uint8)
    for src in (cvops.gaussian_blur5(noise), noise):
        e = cvops.canny(src, 50, 150)
        assert np.array_equal(host(eng.canny(dev(src), 50, 150)), e)
        assert np.array_equal(host(eng.dilate3(dev(e))), cvops.dilate3(e))
        assert np.array_equal(host(eng.dilate3(dev(src))), cvops.dilate3(src))


def test_hysteresis_long_chains(eng):
    """A weak-only snake that crosses every strip boundary with one strong seed at its end."""
    h, w = 200, 330
    g = np.zeros((h, w), np.uint8)
    # serpentine ramp: gentle gradient everywhere (weak), one sharp step (strong)
    yy, xx = np.mgrid[0:h, 0:w]
    g[:] = ((xx * 3 + ((yy // 9) % 2) * 40) % 256).astype(np.uint8)
    g[100:, 200:] = 255
    e = cvops.canny(g, 20, 250)
    assert np.array_equal(host(eng.canny(dev(g), 20, 250)), e)
    for low, high in ((10, 400), (5, 30), (0, 0), (100, 100)):
        assert np.array_equal(host(eng.canny(dev(g), low, high)), cvops.canny(g, low, high))


def _spiral(h, w):
    """A one-pixel-wide weak spiral with a single strong seed at its centre: the longest possible
    propagation chain, crossing every strip boundary many times in both directions."""
    weak = np.zeros((h, w), bool)
    y0, x0, y1, x1 = 1, 1, h - 2, w - 2
    while y1 - y0 >= 4 and x1 - x0 >= 4:
        weak[y0, x0:x1 + 1] = True
        weak[y0:y1 + 1, x1] = True
        weak[y1, x0 + 2:x1 + 1] = True
        weak[y0 + 2:y1 + 1, x0 + 2] = True
        weak[y0 + 2, x0 + 2:x0 + 4] = True
        y0 += 2; x0 += 2; y1 -= 2; x1 -= 2
    return weak


@pytest.mark.parametrize("path", ["cluster", "strips"])
@pytest.mark.parametrize("shape", [(64, 96), (203, 330), (9, 40), (1080, 1920)])
def test_hysteresis_paths_on_synthetic_planes(eng, llfe, path, shape, monkeypatch):
    """Drive the hysteresis stage directly through the plane inputs of both schedules (cluster kernel
    and strip kernels) via llfe_canny on images built to have adversarial weak/strong structure."""
    eng.ctx.set_option("hyst_strips", 1 if path == "strips" else 0)
    try:
        _hysteresis_paths(eng, path, shape)
    finally:
        eng.ctx.set_option("hyst_strips", 0)


def _hysteresis_paths(eng, path, shape):
    h, w = shape
    r = np.random.default_rng(h * 7 + w)
    # (1) blurred noise: dense, branching components; (2) a ramp image whose weak set is a long serpentine
    src = cvops.gaussian_blur5(r.integers(0, 256, shape, dtype=np.uint8))
    for low, high in ((50, 150), (20, 300), (5, 600)):
        if h * w > 500000 and (low, high) != (20, 300):
            continue
        ref = cvops.canny(src, low, high) if h * w <= 500000 else __import__("cv2").Canny(src, low, high)
        assert np.array_equal(host(eng.canny(dev(src), low, high)), ref), (path, shape, low, high)
    yy, xx = np.mgrid[0:h, 0:w]
    g = ((xx * 3 + ((yy // 9) % 2) * 40) % 256).astype(np.uint8)
    g[h // 2:, (2 * w) // 3:] = 255
    ref = cvops.canny(g, 20, 250) if h * w <= 500000 else __import__("cv2").Canny(g, 20, 250)
    assert np.array_equal(host(eng.canny(dev(g), 20, 250)), ref)
    # shape-mask entry point (fused front + hysteresis + dilate) on the same schedule
    bgr = np.stack([g, src, g], axis=-1)
    ref_mask = cvops.shape_mask(bgr) if h * w <= 500000 else None
    if ref_mask is not None:
        assert np.array_equal(host(eng.shape_mask(dev(bgr))), ref_mask)
    # the hysteresis stage alone on adversarial planes: a one-pixel spiral with one strong seed
    weak = _spiral(h, w)
    for seed_at in ("centre", "outer", "none"):
        if path == "strips" and h * w > 500000 and seed_at != "none":
            continue  # the fallback schedule walks a megapixel spiral one row per iteration: correct but slow
        strong = np.zeros_like(weak)
        ys, xs = np.nonzero(weak)
        if seed_at == "centre":
            strong[ys[-1], xs[-1]] = True
        elif seed_at == "outer":
            strong[ys[0], xs[0]] = True
        want = weak & bool(strong.any())         # the spiral is one 8-connected component
        got = host(eng.hysteresis(dev(weak.astype(np.uint8) * 255), dev(strong.astype(np.uint8) * 255)))
        assert np.array_equal(got, want.astype(np.uint8) * 255), (path, shape, seed_at)
        got_d = host(eng.hysteresis(dev(weak.astype(np.uint8)), dev(strong.astype(np.uint8)), dilate=True))
        assert np.array_equal(got_d, cvops.dilate3(want.astype(np.uint8) * 255))
    # random sparse planes against the oracle's flood fill
    wk = r.random(shape) < 0.35
    st = wk & (r.random(shape) < 0.01)
    if h * w <= 500000:
        want = cvops.hysteresis(wk, st)
        got = host(eng.hysteresis(dev(wk.astype(np.uint8)), dev(st.astype(np.uint8))))
        assert np.array_equal(got != 0, want != 0)


@pytest.mark.parametrize("name", ["design_270x480_s1", "design_360x640_s2", "noise_96x160_s3", "design_101x203_s4"])
def test_golden_masks(eng, golden, golden_inputs, name):
    meta, arrays = golden
    img = golden_inputs[name]
    d = dev(img)
    assert np.array_equal(host(eng.shape_mask(d)), arrays[name + "/shape_mask"])
    mask, sums, blurred = eng.shadow_mask(d, want_blurred=True)
    assert np.array_equal(host(blurred), arrays[name + "/shadow_blurred"])
    # any width: OpenCV's vector body AND its tail columns (w % 8 != 0) are reproduced bit for bit
    assert np.array_equal(host(mask), arrays[name + "/shadow_thresh"])
    s, n = [int(v) for v in host(sums)]
    assert cvops.shadow_level(s, n) == meta["cases"][name]["shadow_level"]
    assert np.array_equal(host(eng.font_mask(d)), arrays[name + "/font_mask"])
    if img.shape[0] >= 30 and img.shape[1] >= 100:
        m, t = eng.text_mask(d)
        assert np.array_equal(host(m), arrays[name + "/text_mask"])


@pytest.mark.parametrize("shape", [(64, 96), (53, 40), (11, 16), (135, 256), (30, 1920), (45, 77), (33, 9), (20, 203),
                                   (17, 1001), (12, 1366), (9, 1444), (7, 1921), (40, 67), (40, 69), (2, 13)])
def test_adaptive(eng, shape):
    r = np.random.default_rng(shape[0])
    g = cvops.gaussian_blur5(r.integers(0, 256, shape, dtype=np.uint8))
    mask, sums = eng.adaptive_threshold(dev(g), 2, with_sums=True)
    ref = cvops.adaptive_threshold(g, 2)
    assert np.array_equal(host(mask), ref)
    s, n = [int(v) for v in host(sums)]
    assert n == int((ref == 255).sum()) and s == int(g[ref == 255].astype(np.int64).sum())
    import cv2
    assert np.array_equal(ref, cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, 2))


@pytest.mark.parametrize("seed", range(6))
def test_otsu_text(eng, seed):
    r = np.random.default_rng(seed)
    if seed == 0:
        g = np.full((40, 120), 77, np.uint8)
    elif seed == 1:
        g = cvops.bgr2gray(design_image(120, 200, seed))
    else:
        g = np.clip(r.normal(r.integers(60, 200), 40, (90, 131)), 0, 255).astype(np.uint8)
    t, b = cvops.otsu_binary(g)
    m, thr = eng.otsu(dev(g))
    assert int(thr) == t and np.array_equal(host(m), b)
    img = np.stack([g, g, g], -1)
    m2, _ = eng.text_mask(dev(img))
    assert np.array_equal(host(m2), refpath.text_mask(img))
    light = np.full((40, 120, 3), 250, np.uint8)
    light[5:9, 10:90] = 20     # dark text on white -> Otsu mask mostly white -> inverted
    m3, _ = eng.text_mask(dev(light))
    assert np.array_equal(host(m3), refpath.text_mask(light))


@pytest.mark.parametrize("case", [((216, 384, 3), (200, 112)), ((120, 240, 3), (200, 100)), ((150, 210), (200, 142)),
                                  ((128, 128, 3), (64, 64)), ((128, 256, 3), (64, 32)), ((96, 96), (12, 12)),
                                  ((90, 120, 3), (40, 30)), ((100, 300, 3), (299, 99)), ((64, 64, 3), (64, 64))])
def test_resize_area(eng, case):
    shp, (dw, dh) = case
    src = np.random.default_rng(dw).integers(0, 256, shp, dtype=np.uint8)
    assert np.array_equal(host(eng.resize_area(dev(src), dh, dw)), cvops.resize_area(src, dw, dh))


@pytest.mark.parametrize("case", [((1080, 1920, 3), (1000, 562)), ((2160, 3840, 3), (1000, 562)), ((300, 500, 3), (250, 150)),
                                  ((1200, 1000), (833, 1000)), ((77, 131, 3), (60, 30)), ((64, 64, 3), (32, 32)),
                                  ((40, 60, 3), (90, 70)), ((33, 47), (47, 33)), ((50, 50, 3), (50, 50))])
def test_resize_linear(eng, case):
    import cv2

    shp, (dw, dh) = case
    src = np.random.default_rng(dw + dh).integers(0, 256, shp, dtype=np.uint8)
    got = host(eng.resize_linear(dev(src), dh, dw))
    assert np.array_equal(got, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))
    if src.size <= 500000:
        assert np.array_equal(got, cvops.resize_linear(src, dw, dh))
    batch = np.stack([src, src[::-1].copy()])
    gb = host(eng.resize_linear(dev(batch), dh, dw))
    assert np.array_equal(gb[0], got) and np.array_equal(gb[1], cv2.resize(batch[1], (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("case", [((2250, 4500, 3), (4000, 2000)), ((300, 500, 3), (250, 150)), ((1200, 1000), (833, 1000)),
                                  ((77, 131, 3), (60, 30)), ((64, 64, 3), (32, 32)), ((40, 60, 3), (90, 70)),
                                  ((33, 47), (47, 33)), ((50, 50, 3), (50, 50)), ((9, 7, 3), (5, 4))])
def test_resize_lanczos4(eng, case):
    import cv2

    shp, (dw, dh) = case
    src = np.random.default_rng(dw * 3 + dh).integers(0, 256, shp, dtype=np.uint8)
    got = host(eng.resize_lanczos4(dev(src), dh, dw))
    assert np.array_equal(got, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4))
    if src.size <= 500000:
        assert np.array_equal(got, cvops.resize_lanczos4(src, dw, dh))


def test_golden_resize_and_transform(eng, golden):
    import hashlib
    meta, arrays = golden
    srcs = {"auto_600x2400_s5": design_image(600, 2400, 5), "fit_1300x2200_s6": design_image(1300, 2200, 6)}
    src3 = design_image(512, 768, 7)
    for r in meta["resize"]:
        src = srcs.get(r["name"], src3)
        dh, dw = r["out_shape"][:2]
        out = host(eng.resize_area(dev(src), dh, dw))
        assert hashlib.sha256(out.tobytes()).hexdigest() == r["out_sha256"], r["name"]
    src = design_image(120, 200, 8)
    assert np.array_equal(host(eng.convert_scale_abs(dev(src), 1.2, 0.9)), arrays["transform/bc_1.2_0.9"])
    assert np.array_equal(host(eng.convert_scale_abs(dev(src), 1.3, 1.5)), arrays["transform/bc_1.3_1.5"])
    assert np.array_equal(host(eng.gaussian_blur5(dev(src))), arrays["transform/gaussian_blur"])
    for a in (1.2, 0.9, 1.5, 0.5, 1.3, 1.0, 2.5):
        assert np.array_equal(host(eng.convert_scale_abs(dev(src), a)), cvops.convert_scale_abs(src, a))


def test_full_size_1080p_properties(eng):
    """BASELINE full size: bit-exact against cv2 itself (fast enough on the host) and idempotence of dilate∘mask."""
    img = design_image(1080, 1920, 0)
    m = host(eng.shape_mask(dev(img)))
    assert np.array_equal(m, refpath.shape_mask(img))
    mask, sums = eng.shadow_mask(dev(img))
    assert np.array_equal(host(mask), refpath.shadow_mask(img)[1])
    nz = noise_image(1080, 1920, 1)
    assert np.array_equal(host(eng.shape_mask(dev(nz))), refpath.shape_mask(nz))
    assert np.array_equal(host(eng.shadow_mask(dev(nz))[0]), refpath.shadow_mask(nz)[1])
