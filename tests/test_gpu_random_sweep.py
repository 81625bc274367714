"""Seeded random sweep over image shapes: every mask / colour entry point and the fused pipeline against the
oracle, bit for bit, on shapes nobody picked by hand (odd widths, 1-pixel rows, widths just around the
multiples of 8 and 32 where the kernels switch paths, small batches)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cvops, refpath  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image, noise_image  # noqa: E402


def _shapes(n, seed):
    r = np.random.default_rng(seed)
    out = []
    for i in range(n):
        h = int(r.integers(1, 200))
        w = int(r.integers(1, 330))
        if i % 3 == 0:
            w = max(8, (w // 8) * 8)          # fused-kernel widths
        if i % 7 == 0:
            w = int(r.choice([31, 32, 33, 63, 64, 65, 255, 256, 257]))
        out.append((h, w))
    return out


@pytest.fixture(scope="module")
def eng():
    import low_level_feature_extraction_b200 as pkg
    return pkg.engine(0)


def _image(h, w, seed):
    if seed % 2 and h >= 16 and w >= 16:
        return design_image(h, w, seed)
    return noise_image(h, w, seed)


@pytest.mark.parametrize("shape", _shapes(40, 2024))
def test_masks_on_random_shapes(eng, shape):
    h, w = shape
    seed = h * 1000 + w
    nb = 1 + seed % 3
    batch = np.stack([_image(h, w, seed + i) for i in range(nb)])
    d = torch.from_numpy(batch).cuda()
    if h >= 2 and w >= 2:
        sm = eng.shape_mask(d).cpu().numpy()
    mask, sums = eng.shadow_mask(d)
    fm = eng.font_mask(d).cpu().numpy()
    mask, sums = mask.cpu().numpy(), sums.cpu().numpy()
    for i in range(nb):
        if h >= 2 and w >= 2:
            assert np.array_equal(sm[i], cvops.shape_mask(batch[i]))
        _, m_ref, s_ref, n_ref, _ = cvops.shadow_parts(batch[i])
        assert np.array_equal(mask[i], m_ref)
        assert (int(sums[i, 0]), int(sums[i, 1])) == (s_ref, n_ref)
        assert np.array_equal(fm[i], cvops.font_mask(batch[i]))
    if h >= 30 and w >= 100:
        tm, _ = eng.text_mask(d)
        for i in range(nb):
            assert np.array_equal(tm[i].cpu().numpy(), refpath.text_mask(batch[i]))


@pytest.mark.parametrize("shape", _shapes(24, 77))
def test_colours_and_fused_pipeline_on_random_shapes(eng, shape):
    h, w = shape
    seed = h * 1000 + w
    img = _image(h, w, seed)
    noise = cvops.make_noise((h * w, 3), seed).reshape(h, w, 3)
    d, dn = torch.from_numpy(img[None]).cuda(), torch.from_numpy(noise[None]).cuda()
    ref = cvops.unique_colors(cvops.apply_noise(img.reshape(-1, 3)[:, ::-1], noise.reshape(-1, 3)))
    ref_keys = (ref[:, 0].astype(np.int64) << 16) | (ref[:, 1].astype(np.int64) << 8) | ref[:, 2]
    keys, count = eng.unique_colors(d, noise=dn, max_unique=1 << 17)
    n = int(count[0])
    assert n == len(ref_keys) and np.array_equal(keys[0, :n].cpu().numpy().astype(np.int64) & 0xFFFFFF, ref_keys)
    if h >= 2 and w >= 2:
        out = eng.pipeline(d, noise=dn, max_unique=1 << 17)
        n2 = int(out["count"][0])
        assert n2 == len(ref_keys)
        assert np.array_equal(out["keys"][0, :n2].cpu().numpy().astype(np.int64) & 0xFFFFFF, ref_keys)
        assert np.array_equal(out["shape_mask"][0].cpu().numpy(), cvops.shape_mask(img))
        _, m_ref, s_ref, n_ref, _ = cvops.shadow_parts(img)
        assert np.array_equal(out["shadow_mask"][0].cpu().numpy(), m_ref)
        assert [int(v) for v in out["shadow_sums"][0].cpu()] == [s_ref, n_ref]


def _colour_list(kind, u, r):
    """u distinct RGB colours (sorted like np.unique): uniform, a few tight blobs, or a gray-ish ramp."""
    if kind == "uniform":
        px = r.integers(0, 256, (u * 2 + 8, 3))
    elif kind == "blobs":
        c = r.integers(20, 236, (int(r.integers(2, 9)), 3))
        px = c[r.integers(0, len(c), u * 3 + 8)] + r.integers(-18, 19, (u * 3 + 8, 3))
    else:
        t = r.integers(0, 256, (u * 3 + 8, 1))
        px = t + r.integers(-6, 7, (u * 3 + 8, 3))
    uq = np.unique(np.clip(px, 0, 255).astype(np.uint8), axis=0)
    if len(uq) > u:
        uq = uq[np.sort(r.choice(len(uq), u, replace=False))]
    return uq


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 16])
def test_kmeans_batch_of_mixed_sizes_matches_cv2(eng, k):
    """One launch over lists of very different lengths (every tier of the k-means kernel, from a single colour to
    a list that needs the global-memory variant) against cv2.kmeans itself: labels and float32 centres bit for bit."""
    import cv2

    r = np.random.default_rng(1000 + k)
    sizes = [1, 2, 5, 17, 300, 2500, 9000, 16000, 24000, 40000, 60000]
    kinds = ["uniform", "blobs", "ramp"]
    lists = [_colour_list(kinds[(i + k) % 3], u, r) for i, u in enumerate(sizes)]
    mu = 1 << 16
    keys = np.zeros((len(lists), mu), np.int32)
    cnt = np.zeros(len(lists), np.int32)
    for i, uq in enumerate(lists):
        keys[i, :len(uq)] = (uq[:, 0].astype(np.int64) << 16 | uq[:, 1].astype(np.int64) << 8 | uq[:, 2]).astype(np.int32)
        cnt[i] = len(uq)
    seeds = [int(s) for s in r.integers(0, 1 << 31, len(lists))]
    centers, labels, comp, kused = eng.kmeans_unique(torch.from_numpy(keys).cuda(), torch.from_numpy(cnt).cuda(), k, seeds)
    for i, uq in enumerate(lists):
        ka = min(k, len(uq))
        assert int(kused[i]) == ka
        if ka <= 1:       # color_extractor.py:183-186: no cv2.kmeans call at all
            continue
        cv2.setRNGSeed(seeds[i])
        c_ref, l_ref, ce_ref = cv2.kmeans(np.float32(uq), ka, None, refpath.KMEANS_CRITERIA, 10, cv2.KMEANS_PP_CENTERS)
        assert np.array_equal(labels[i, :len(uq)].cpu().numpy(), l_ref.ravel()), (k, len(uq))
        assert np.array_equal(centers[i, :ka].cpu().numpy(), ce_ref), (k, len(uq))
        assert abs(float(comp[i]) - c_ref) <= 1e-9 * max(1.0, c_ref)
