"""Per-pixel k-means of one image (BASELINE config 5) through the C ABI on the GPU.

World size 1 here (one GPU per gpurun box by default); the multi-rank host logic is covered
by tests/test_dist_cpu.py with gloo, and `test_two_shards_accumulate_like_one` checks that
the device accumulator is shard-invariant (what makes the all-reduced result bit-identical).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from low_level_feature_extraction_b200.synth import design_image, noise_image  # noqa: E402
from oracle import cvops, refpath  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import low_level_feature_extraction_b200 as pkg
    return pkg.engine(0)


def _init_from_pixels(px, k, seed):
    return np.float32(px[np.random.default_rng(seed).choice(len(px), k, replace=False)])


@pytest.mark.parametrize("case", [("design", 96, 128, 16, 42), ("design", 45, 77, 5, 1), ("noise", 64, 96, 8, 3)])
def test_pixel_kmeans_matches_exact_oracle(eng, case):
    from low_level_feature_extraction_b200.dist import PixelKMeans

    kind, h, w, k, seed = case
    img = design_image(h, w, seed) if kind == "design" else noise_image(h, w, seed)
    px = img.reshape(-1, 3)[:, ::-1]
    init = _init_from_pixels(px, k, seed)
    c_ref, l_ref, it_ref, s_ref, n_ref = cvops.lloyd_exact(px, init)
    res = PixelKMeans(eng).fit(torch.from_numpy(img).cuda(), torch.from_numpy(init), want_labels=True)
    assert res.iters == it_ref
    assert np.array_equal(res.centers.cpu().numpy(), c_ref)            # bit-exact centres
    assert np.array_equal(res.labels.cpu().numpy(), l_ref.astype(np.uint8))  # bit-exact pixel labels
    s = res.sums_counts.cpu().numpy()
    assert np.array_equal(s[:, :3], s_ref) and np.array_equal(s[:, 3], n_ref)
    # north-star tolerance against raw cv2.kmeans from the same seeded centroids: <= 1e-3 relative on centroids
    if kind == "design":
        c_cv, _, _ = refpath.kmeans_pixels(np.float32(px), init)
        assert np.abs(res.centers.cpu().numpy() - c_cv).max() / 255.0 <= 1e-3


def test_pixel_kmeans_empty_cluster_repair(eng):
    from low_level_feature_extraction_b200.dist import PixelKMeans

    img = design_image(40, 56, 9)
    px = img.reshape(-1, 3)[:, ::-1]
    init = _init_from_pixels(px, 5, 7)
    init[3] = init[1]
    init[4] = init[1]          # two empty clusters on the first update, same donor twice
    c_ref, l_ref, it_ref, s_ref, n_ref = cvops.lloyd_exact(px, init)
    res = PixelKMeans(eng).fit(torch.from_numpy(img).cuda(), torch.from_numpy(init), want_labels=True)
    assert res.iters == it_ref
    assert np.array_equal(res.centers.cpu().numpy(), c_ref)
    assert np.array_equal(res.labels.cpu().numpy(), l_ref.astype(np.uint8))


def test_two_shards_accumulate_like_one(eng):
    img = design_image(101, 203, 4)
    px = img.reshape(-1, 3)[:, ::-1]
    init = torch.from_numpy(_init_from_pixels(px, 16, 0)).cuda()
    d = torch.from_numpy(img).cuda()
    whole = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    eng.kmeans_pixels_step(d, init, whole)
    parts = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    for r0, r1 in ((0, 37), (37, 101)):
        eng.kmeans_pixels_step(d[r0:r1].contiguous(), init, parts)
    assert torch.equal(whole, parts)
    lab, _ = cvops.assign(px.astype(np.float32), init.cpu().numpy())
    assert np.array_equal(whole[:, 3].cpu().numpy(), np.bincount(lab, minlength=16))


def test_full_size_step_properties(eng):
    """At a config-5-like row-shard size the oracle is too slow: check conservation laws instead."""
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = torch.randint(0, 256, (2048, 4096, 3), dtype=torch.uint8, device="cuda", generator=g)
    init = torch.rand((16, 3), device="cuda", generator=g) * 255
    sums = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    lab = torch.empty((2048 * 4096,), dtype=torch.uint8, device="cuda")
    eng.kmeans_pixels_step(rows, init, sums, lab)
    assert int(sums[:, 3].sum()) == 2048 * 4096
    tot = rows.reshape(-1, 3).to(torch.int64).sum(0).flip(0)           # RGB totals
    assert torch.equal(sums[:, :3].sum(0), tot)
    assert torch.equal(torch.bincount(lab.to(torch.int64), minlength=16), sums[:, 3])
