"""The fused front kernel (one read of the image -> edge planes + shadow mask + colour bitmap) vs the oracle,
at the shapes that exercise its halo lanes, column bands, row bands and image borders."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cvops, refpath  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image, noise_image  # noqa: E402

# (h, w): w % 8 == 0 -> fused path.  240 / 248 / 256: one-warp boundary; 960 / 968: CTA column band boundary;
# 1928, 2000: two bands with a ragged last warp; tiny heights: reflect-101 with h < 5; 1080 rows: 8 row bands.
SHAPES = [(1, 8), (2, 16), (3, 24), (5, 8), (7, 40), (33, 240), (33, 248), (20, 256), (50, 960), (50, 968), (9, 1928),
          (12, 2000), (97, 64), (200, 104), (383, 136), (384, 8), (385, 16)]


@pytest.fixture(scope="module")
def eng():
    import low_level_feature_extraction_b200 as pkg
    return pkg.engine(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(params=["split", "inline"])
def shadow_path(request, eng):
    """The adaptive threshold in its own kernel (k_shadow, default) or inside the fused front kernel."""
    eng.ctx.set_option("shadow_inline", 1 if request.param == "inline" else 0)
    yield request.param
    eng.ctx.set_option("shadow_inline", 0)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["noise", "design"])
def test_fused_pipeline_matches_oracle(eng, shape, kind, shadow_path):
    h, w = shape
    img = noise_image(h, w, h * 7 + w) if kind == "noise" else design_image(h, w, h + w)
    noise = cvops.make_noise((h * w, 3), 7 + h).reshape(h, w, 3)
    out = eng.pipeline(dev(img[None]), noise=dev(noise[None]), max_unique=1 << 15)
    if h >= 2:
        assert np.array_equal(out["shape_mask"][0].cpu().numpy(), cvops.shape_mask(img))
    _, m, s, n, _ = cvops.shadow_parts(img)
    assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), m)
    assert out["shadow_sums"][0].cpu().tolist() == [s, n]
    px = cvops.apply_noise(cvops.bgr2rgb(img).reshape(-1, 3), noise.reshape(-1, 3))
    u = cvops.unique_colors(px)
    c = int(out["count"][0])
    assert c == len(u)
    k = out["keys"][0].cpu().numpy()[:min(c, 1 << 15)].astype(np.uint32)
    got = np.stack([k >> 16, (k >> 8) & 255, k & 255], 1).astype(np.uint8)
    assert np.array_equal(got, u[:len(got)])


def test_fused_equals_unfused_with_device_noise(eng):
    """Device noise is a function of (seed, pixel index): the fused and unfused colour passes must agree."""
    batch = np.stack([design_image(135, 240, s) for s in range(3)])
    d = dev(batch)
    fused = eng.pipeline(d, seed=77, max_unique=1 << 15)
    keys, count = eng.unique_colors(d, None, seed=77, max_unique=1 << 15)
    assert np.array_equal(fused["count"].cpu().numpy(), count.cpu().numpy())
    for i in range(3):
        c = int(count[i])
        assert np.array_equal(fused["keys"][i, :c].cpu().numpy(), keys[i, :c].cpu().numpy())


def test_fused_thresholds_and_partial_outputs(eng):
    img = noise_image(90, 160, 3)
    for low, high in ((0, 0), (20, 250), (100, 100), (400, 900), (3000, 4000)):
        got = eng.shape_mask(dev(img), low, high).cpu().numpy()
        assert np.array_equal(got, cvops.shape_mask(img, low, high)), (low, high)
    only_shadow = eng.pipeline(dev(img[None]), shapes=False, colors=False)
    assert np.array_equal(only_shadow["shadow_mask"][0].cpu().numpy(), cvops.shadow_parts(img)[1])
    only_colors = eng.pipeline(dev(img[None]), shapes=False, shadows=False, noise=dev(np.zeros((1,) + img.shape, np.int8)))
    assert int(only_colors["count"][0]) == len(np.unique(img.reshape(-1, 3), axis=0))


def test_fused_full_size_vs_cv2(eng, shadow_path):
    for img in (design_image(1080, 1920, 3), noise_image(1080, 1920, 4), design_image(1125, 2000, 5)):
        out = eng.pipeline(dev(img[None]), colors=False)
        assert np.array_equal(out["shape_mask"][0].cpu().numpy(), refpath.shape_mask(img))
        assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), refpath.shadow_mask(img)[1])


def test_4k_and_tall_images_vs_cv2(eng):
    """BASELINE config 3 size (3840x2160: 120 plane words per row -> 4 words per lane, 16-CTA clusters),
    a tall narrow image (many rows per strip), and a batch of two so that image offsets are exercised."""
    imgs = [design_image(2160, 3840, 7), noise_image(2160, 3840, 8)]
    batch = np.stack(imgs)
    out = eng.pipeline(dev(batch), colors=False)
    for i, img in enumerate(imgs):
        assert np.array_equal(out["shape_mask"][i].cpu().numpy(), refpath.shape_mask(img)), i
        ref_blur, ref_mask = refpath.shadow_mask(img)
        assert np.array_equal(out["shadow_mask"][i].cpu().numpy(), ref_mask), i
        s, n = out["shadow_sums"][i].cpu().tolist()
        assert n == int((ref_mask == 255).sum()) and s == int(ref_blur[ref_mask == 255].astype(np.int64).sum())
    tall = design_image(1900, 264, 9)
    out = eng.pipeline(dev(tall[None]), colors=False)
    assert np.array_equal(out["shape_mask"][0].cpu().numpy(), refpath.shape_mask(tall))
    # taller than the cluster kernel's per-warp row window: takes the strip kernels
    taller = noise_image(4000, 72, 10)
    out = eng.pipeline(dev(taller[None]), colors=False)
    assert np.array_equal(out["shape_mask"][0].cpu().numpy(), refpath.shape_mask(taller))
    assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), refpath.shadow_mask(taller)[1])


def test_4k_palette_k16(eng):
    """Config 3's palette: k = 16 on a 4K image, injected noise, against cv2.kmeans on the same unique colours."""
    import cv2

    img = design_image(2160, 3840, 11)
    noise = cvops.make_noise((2160 * 3840, 3), 77).reshape(2160, 3840, 3)
    keys, count = eng.unique_colors(dev(img), dev(noise), max_unique=1 << 17)
    px = cvops.apply_noise(cvops.bgr2rgb(img).reshape(-1, 3), noise.reshape(-1, 3))
    u = cvops.unique_colors(px)
    c = int(count)
    k = keys[:c].cpu().numpy().astype(np.uint32)
    assert c == len(u) and np.array_equal(np.stack([(k >> 16) & 255, (k >> 8) & 255, k & 255], 1).astype(np.uint8), u)
    centers, labels, comp, kused = eng.kmeans_unique(keys, count, 16, 4242)
    cv2.setRNGSeed(4242)
    comp_cv, lab_cv, cen_cv = cv2.kmeans(np.float32(u), 16, None, (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 200, 0.2),
                                          10, cv2.KMEANS_PP_CENTERS)
    assert np.array_equal(centers[0].cpu().numpy(), cen_cv)
    assert np.array_equal(labels[0, :c].cpu().numpy(), lab_cv.reshape(-1))
    assert abs(float(comp[0]) - comp_cv) <= 1e-9 * max(1.0, comp_cv)


def test_unfused_path_still_matches(eng):
    """The "unfused" option forces the per-stage kernels (the path odd widths take)."""
    img = design_image(96, 160, 1)
    eng.ctx.set_option("unfused", 1)
    try:
        out = eng.pipeline(dev(img[None]), colors=False)
        assert np.array_equal(out["shape_mask"][0].cpu().numpy(), cvops.shape_mask(img))
        assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), cvops.shadow_parts(img)[1])
    finally:
        eng.ctx.set_option("unfused", 0)


def test_tall_image_single_column_of_bands(eng, shadow_path):
    """A tall, narrow image (few, very tall row bands per CTA column): the masked sum / count accumulated per lane over
    thousands of rows must still be exact, also when the adaptive mask fires on very many pixels."""
    h, w = 3000, 64
    img = noise_image(h, w, 5)
    img[::2] = 255                      # rows alternate white / noise: the adaptive mask fires on very many pixels
    d = torch.from_numpy(img[None]).cuda()
    out = eng.pipeline(d, colors=False)
    _, m_ref, s_ref, n_ref, _ = cvops.shadow_parts(img)
    assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), m_ref)
    assert [int(v) for v in out["shadow_sums"][0].cpu()] == [s_ref, n_ref]
    assert np.array_equal(out["shape_mask"][0].cpu().numpy(), cvops.shape_mask(img))


def test_results_do_not_depend_on_the_chunking(eng):
    """llfe_analyze launches each front kernel once per super-chunk by default; the device noise, the masks and the
    palettes must be the same for any images-per-launch setting (the schedule is not part of the result)."""
    batch = np.stack([design_image(135, 240, s) for s in range(9)])
    d = dev(batch)
    ref = None
    try:
        for chunk in (256, 4, 1):
            eng.ctx.set_option("chunk", chunk)
            out = eng.pipeline(d, seed=5, max_unique=1 << 15)
            got = {k: out[k].cpu().numpy().copy() for k in ("shape_mask", "shadow_mask", "shadow_sums", "count")}
            got["keys"] = [out["keys"][i, :int(out["count"][i])].cpu().numpy().copy() for i in range(9)]
            if ref is None:
                ref = got
                continue
            for k in ("shape_mask", "shadow_mask", "shadow_sums", "count"):
                assert np.array_equal(got[k], ref[k]), (chunk, k)
            for i in range(9):
                assert np.array_equal(got["keys"][i], ref["keys"][i]), (chunk, i)
    finally:
        eng.ctx.set_option("chunk", 256)
