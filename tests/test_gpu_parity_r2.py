"""Round-2 parity holes, closed: truncated colour lists, cv2's maxCount clamp, the n_colors <= 1 return of
color_extractor.py:185-186, contexts used from other threads / with another device current."""
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import cv2  # noqa: E402
from oracle import cvops  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    import low_level_feature_extraction_b200 as pkg
    return pkg.engine(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def photo_like(h, w, seed):
    """Smooth ramps + per-channel noise: far more than 65 536 distinct colours at 270x480 and up."""
    r = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], -1).astype(np.int32)
    img += r.integers(-12, 13, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_truncated_list_is_never_clustered(eng):
    """count > max_unique: k_used = -1 + status bit, no palette from a truncated list."""
    img = photo_like(300, 500, 1)
    zero = np.zeros(img.shape, np.int8)
    keys, count = eng.unique_colors(dev(img), dev(zero), max_unique=4096)
    u = len(np.unique(img.reshape(-1, 3), axis=0))
    assert int(count) == u > 4096
    centers, labels, comp, kused = eng.kmeans_unique(keys, count, 5, 7)
    assert int(kused[0]) == -1 and int(eng.last_status[0]) & 2
    assert float(centers.abs().sum()) == 0.0


def test_batch_overflow_is_redone_with_a_full_list(eng):
    """A batch mixing design frames and photo-like frames: the photo-like ones overflow the batched list and are
    redone alone; every palette equals the one from a batch whose list holds every colour, on device and through the
    host-buffer path (same device noise in all three)."""
    from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig

    h, w = 270, 480
    imgs = np.stack([design_image(h, w, 1), photo_like(h, w, 2), design_image(h, w, 3), photo_like(h, w, 4)])
    d = dev(imgs)
    small = BatchAnalyzer(0, h, w, BatchConfig(max_unique=1 << 14, seed=5))
    big = BatchAnalyzer(0, h, w, BatchConfig(max_unique=1 << 17, seed=5))
    a = small.run_device(d)
    b = big.run_device(d)
    cnt = b["count"].cpu().numpy()
    assert (cnt[[1, 3]] > (1 << 14)).all() and (cnt[[0, 2]] <= (1 << 14)).all() and (cnt < (1 << 17)).all()
    for k in ("count", "k_used", "centers", "cluster_sizes"):
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy()), k
    assert (a["k_used"].cpu().numpy() == 5).all()
    # resolve=False leaves the sentinel for the caller
    c = small.run_device(d, resolve=False)
    assert c["k_used"].cpu().numpy().tolist() == [5, -1, 5, -1]
    assert small.resolve_overflow(d, c) == [1, 3]
    assert np.array_equal(c["centers"].cpu().numpy(), b["centers"].cpu().numpy())
    # host-buffer path (stages of 4 images: the chunk-local noise indices are the same as on the device path)
    pinned = torch.from_numpy(imgs).pin_memory()
    hs = BatchAnalyzer(0, h, w, BatchConfig(max_unique=1 << 14, seed=5, host_chunk=4, host_streams=1)).run_host(pinned)
    hb = BatchAnalyzer(0, h, w, BatchConfig(max_unique=1 << 17, seed=5, host_chunk=4, host_streams=1)).run_host(pinned)
    for k in ("count", "k_used", "centers", "cluster_sizes"):
        assert np.array_equal(hs[k].numpy(), hb[k].numpy()), k
    assert (hs["k_used"].numpy() == 5).all()


def test_request_batcher_on_a_photo_like_frame_equals_extract_colors():
    """A >65 536-colour frame through RequestBatcher (batched list overflows -> redone) gives the palette
    ColorExtractor computes from the same noised pixels."""
    from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig
    from low_level_feature_extraction_b200.services import ColorExtractor
    from low_level_feature_extraction_b200.services.batching import RequestBatcher

    h, w = 360, 640
    img = photo_like(h, w, 9)
    with RequestBatcher(device=0, max_batch=4, max_wait_ms=5.0) as rb:
        got = rb.analyze(img)["colors"]
    # the same pipeline with injected zero noise is not what the batcher ran (device noise), so compare with the
    # device-noise palette of a batch whose list is long enough, pushed through the reference's palette tail
    an = BatchAnalyzer(0, h, w, BatchConfig(max_unique=1 << 18, host_chunk=4))
    out = an.run_device(dev(img[None]))
    assert int(out["count"][0]) > (1 << 16)
    k = int(out["k_used"][0])
    c8 = out["centers"][0, :k].cpu().numpy().astype(np.uint8)
    want = ColorExtractor._palette_from_clusters(c8, out["cluster_sizes"][0, :k].cpu().numpy().astype(np.int64))
    assert (got.primary, got.background, list(got.accent)) == (want.primary, want.background, list(want.accent))
    assert got.metadata["success"]


@pytest.mark.parametrize("seed", [5])
def test_kmeans_iteration_cap_is_cv2s_100(eng, seed):
    """eps = 0 on a slow-converging list: cv2 stops at 100 iterations whatever maxCount >= 100 says."""
    r = np.random.default_rng(seed)
    data = r.integers(0, 256, (20000, 3)).astype(np.uint8)
    uniq = np.unique(data, axis=0)
    keys = (uniq[:, 0].astype(np.uint32) << 16) | (uniq[:, 1].astype(np.uint32) << 8) | uniq[:, 2]
    cv2.setRNGSeed(7)
    comp_cv, lab_cv, cen_cv = cv2.kmeans(uniq.astype(np.float32), 16, None,
                                         (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 200, 0.0), 1,
                                         cv2.KMEANS_PP_CENTERS)
    centers, labels, comp, kused = eng.kmeans_unique(dev(keys.astype(np.int64).astype(np.int32)[None]),
                                                     dev(np.array([len(keys)], np.int32)), 16, 7, attempts=1,
                                                     max_iter=200, eps=0.0)
    assert np.array_equal(centers[0].cpu().numpy(), cen_cv)
    assert np.array_equal(labels[0, :len(keys)].cpu().numpy(), lab_cv.ravel())
    assert abs(float(comp[0]) - comp_cv) <= 1e-9 * comp_cv


def test_single_cluster_request_returns_the_unique_list():
    """color_extractor.py:185-186: n_colors <= 1 returns ALL unique colours as centres and zero labels."""
    from low_level_feature_extraction_b200.services import ColorExtractor

    img = design_image(60, 90, 3)
    px = img[..., ::-1].reshape(-1, 3)
    uniq = np.unique(px, axis=0)
    for k in (1, 0):
        centers, labels = ColorExtractor._get_dominant_colors(px, k)
        assert np.array_equal(centers, uniq) and labels.shape == (len(uniq),) and not labels.any()
    flat = np.full((20, 30, 3), 77, np.uint8)
    centers, labels = ColorExtractor._get_dominant_colors(flat.reshape(-1, 3), 5)
    assert np.array_equal(centers, [[77, 77, 77]]) and labels.tolist() == [0]
    np.random.seed(3)
    cf = ColorExtractor.extract_colors(img, 1)
    assert cf.metadata["success"] and cf.primary.startswith("#")


def test_context_works_from_another_thread_and_device(llfe):
    """Every entry point makes the context's device current itself: a context created on one thread
    is used from a fresh thread, and -- on a multi-GPU box -- while another device is current."""
    from low_level_feature_extraction_b200.services import ShapeAnalyzer

    img = design_image(64, 96, 1)
    want = cvops.shape_mask(img)
    res = {}

    def worker():
        res["mask"] = ShapeAnalyzer.preprocess_image(img)

    ShapeAnalyzer.preprocess_image(img)      # creates the process-wide context on this thread
    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert np.array_equal(res["mask"], want)
    if torch.cuda.device_count() >= 2:
        ctx1 = llfe.Context(1)
        try:
            mask = np.empty(img.shape[:2], np.uint8)

            def worker1():
                torch.cuda.set_device(0)
                ctx1.call("llfe_shape_mask_host", img, img.shape[0], img.shape[1], 50, 150, mask)
                res["dev"] = torch.cuda.current_device()

            t = threading.Thread(target=worker1)
            t.start()
            t.join()
            assert np.array_equal(mask, want) and res["dev"] == 0
        finally:
            ctx1.close()


def test_debug_buffer_is_validated(eng):
    from low_level_feature_extraction_b200 import LlfeError

    host = np.zeros(64, np.int64)
    with pytest.raises(LlfeError):
        eng.ctx.call("llfe_set_debug_buffer", b"kmeans", host, host.nbytes)
    with pytest.raises(LlfeError):
        eng.ctx.set_option("no_such_option", 1)
    buf = torch.zeros((1, 10, 8), dtype=torch.int64, device="cuda")
    eng.ctx.set_debug_buffer("kmeans", buf)
    eng.ctx.set_debug_buffer("kmeans", None)
