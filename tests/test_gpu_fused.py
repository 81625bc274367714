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


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["noise", "design"])
def test_fused_pipeline_matches_oracle(eng, shape, kind):
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


def test_fused_full_size_vs_cv2(eng):
    for img in (design_image(1080, 1920, 3), noise_image(1080, 1920, 4), design_image(1125, 2000, 5)):
        out = eng.pipeline(dev(img[None]), colors=False)
        assert np.array_equal(out["shape_mask"][0].cpu().numpy(), refpath.shape_mask(img))
        assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), refpath.shadow_mask(img)[1])


def test_unfused_path_still_matches(eng):
    """LLFE_UNFUSED=1 forces the per-stage kernels (the path odd widths take)."""
    img = design_image(96, 160, 1)
    os.environ["LLFE_UNFUSED"] = "1"
    try:
        out = eng.pipeline(dev(img[None]), colors=False)
        assert np.array_equal(out["shape_mask"][0].cpu().numpy(), cvops.shape_mask(img))
        assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), cvops.shadow_parts(img)[1])
    finally:
        del os.environ["LLFE_UNFUSED"]
