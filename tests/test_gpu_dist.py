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


@pytest.mark.parametrize("histogram", [True, "launch-per-iteration", False])
@pytest.mark.parametrize("case", [("design", 96, 128, 16, 42), ("design", 45, 77, 5, 1), ("noise", 64, 96, 8, 3)])
def test_pixel_kmeans_matches_exact_oracle(eng, case, histogram):
    from low_level_feature_extraction_b200.dist import PixelKMeans as _PKM

    def PixelKMeans(e, histogram):       # True: persistent cooperative kernel; "launch-per-iteration": step + update launches
        km = _PKM(e, histogram=bool(histogram))
        km.persistent = histogram is True
        return km

    kind, h, w, k, seed = case
    img = design_image(h, w, seed) if kind == "design" else noise_image(h, w, seed)
    px = img.reshape(-1, 3)[:, ::-1]
    init = _init_from_pixels(px, k, seed)
    c_ref, l_ref, it_ref, s_ref, n_ref = cvops.lloyd_exact(px, init)
    res = PixelKMeans(eng, histogram=histogram).fit(torch.from_numpy(img).cuda(), torch.from_numpy(init),
                                                    want_labels=True)
    assert res.iters == it_ref
    assert np.array_equal(res.centers.cpu().numpy(), c_ref)            # bit-exact centres
    assert np.array_equal(res.labels.cpu().numpy(), l_ref.astype(np.uint8))  # bit-exact pixel labels
    s = res.sums_counts.cpu().numpy()
    assert np.array_equal(s[:, :3], s_ref) and np.array_equal(s[:, 3], n_ref)
    # north-star tolerance against raw cv2.kmeans from the same seeded centroids: <= 1e-3 relative on centroids
    if kind == "design":
        c_cv, _, _ = refpath.kmeans_pixels(np.float32(px), init)
        assert np.abs(res.centers.cpu().numpy() - c_cv).max() / 255.0 <= 1e-3


@pytest.mark.parametrize("histogram", [True, "launch-per-iteration", False])
def test_pixel_kmeans_empty_cluster_repair(eng, histogram):
    from low_level_feature_extraction_b200.dist import PixelKMeans as _PKM

    def PixelKMeans(e, histogram):
        km = _PKM(e, histogram=bool(histogram))
        km.persistent = histogram is True
        return km

    img = design_image(40, 56, 9)
    px = img.reshape(-1, 3)[:, ::-1]
    init = _init_from_pixels(px, 5, 7)
    init[3] = init[1]
    init[4] = init[1]          # two empty clusters on the first update, same donor twice
    c_ref, l_ref, it_ref, s_ref, n_ref = cvops.lloyd_exact(px, init)
    res = PixelKMeans(eng, histogram=histogram).fit(torch.from_numpy(img).cuda(), torch.from_numpy(init),
                                                    want_labels=True)
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


@pytest.mark.parametrize("shape", [(45, 77), (101, 203), (64, 96), (1, 5), (300, 517)])
def test_colour_histogram_equals_numpy_unique(eng, shape):
    """Count table + ordered compaction == np.unique(pixels, axis=0, return_counts=True), for every split into parts."""
    h, w = shape
    img = design_image(h, w, 11) if h * w > 16 else noise_image(h, w, 11)
    d = torch.from_numpy(img).cuda()
    hist = torch.zeros((1 << 24,), dtype=torch.int32, device="cuda")
    eng.pixels_histogram(d[: h // 2].contiguous(), hist)          # two row shards accumulate into one table
    eng.pixels_histogram(d[h // 2:].contiguous(), hist)
    uq, cnt = np.unique(img.reshape(-1, 3)[:, ::-1], axis=0, return_counts=True)   # RGB rows, lexicographic
    key_ref = (uq[:, 0].astype(np.int64) << 16) | (uq[:, 1].astype(np.int64) << 8) | uq[:, 2]
    keys, counts = eng.histogram_compact(hist)
    assert np.array_equal(keys.cpu().numpy(), key_ref) and np.array_equal(counts.cpu().numpy(), cnt)
    for parts in (2, 3, 8):
        got_k, got_c = [], []
        for part in range(parts):
            kk, cc = eng.histogram_compact(hist, part, parts)
            assert np.all((kk.cpu().numpy() // 2048) % parts == part)
            got_k.append(kk.cpu().numpy())
            got_c.append(cc.cpu().numpy())
        order = np.argsort(np.concatenate(got_k), kind="stable")
        assert np.array_equal(np.concatenate(got_k)[order], key_ref)
        assert np.array_equal(np.concatenate(got_c)[order], cnt)


def test_histogram_step_equals_pixel_step(eng):
    """The weighted step over distinct colours adds exactly what the per-pixel step adds; labels agree pixel by pixel."""
    g = torch.Generator(device="cuda").manual_seed(5)
    rows = (torch.randint(0, 40, (1024, 2048, 3), dtype=torch.uint8, device="cuda", generator=g) * 6
            + torch.randint(0, 3, (1024, 2048, 3), dtype=torch.uint8, device="cuda", generator=g))
    rows[:300, :, :] = torch.tensor([250, 3, 77], dtype=torch.uint8, device="cuda")   # one colour with a huge count
    init = torch.rand((16, 3), device="cuda", generator=g) * 255
    a = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    lab_a = torch.empty((1024 * 2048,), dtype=torch.uint8, device="cuda")
    eng.kmeans_pixels_step(rows, init, a, lab_a)
    hist = torch.zeros((1 << 24,), dtype=torch.int32, device="cuda")
    eng.pixels_histogram(rows, hist)
    assert int(hist.sum()) == 1024 * 2048
    keys, counts = eng.histogram_compact(hist)
    b = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    lab_e = torch.empty((keys.numel(),), dtype=torch.uint8, device="cuda")
    eng.kmeans_hist_step(keys, counts, init, b, lab_e)
    assert torch.equal(a, b)
    lut = torch.zeros((1 << 24,), dtype=torch.uint8, device="cuda")
    eng.hist_labels_to_lut(keys, lab_e, lut)
    lab_b = torch.empty_like(lab_a)
    eng.pixels_lookup(rows, lut, lab_b)
    assert torch.equal(lab_a, lab_b)
    # unaligned base pointer (head pixels) and a ragged tail
    flat = rows.reshape(-1)[3 * 5: 3 * 5 + 3 * 100003].clone()
    sub = torch.empty((flat.numel() + 3,), dtype=torch.uint8, device="cuda")[3:]
    sub.copy_(flat)
    h2 = torch.zeros((1 << 24,), dtype=torch.int32, device="cuda")
    eng.pixels_histogram(sub.view(1, -1, 3), h2)
    ref = torch.bincount((sub.view(-1, 3)[:, 2].to(torch.int64) << 16) | (sub.view(-1, 3)[:, 1].to(torch.int64) << 8)
                         | sub.view(-1, 3)[:, 0].to(torch.int64), minlength=1 << 24)
    assert torch.equal(h2.to(torch.int64), ref)
    l2 = torch.empty((100003,), dtype=torch.uint8, device="cuda")
    eng.pixels_lookup(sub.view(1, -1, 3), lut, l2)
    assert torch.equal(l2, lab_a[5: 5 + 100003])


def test_config5_full_size_properties(eng):
    """BASELINE config 5 at its real size (one 16384 x 16384 image = 2^28 pixels, K = 16), where the oracle is far too
    slow: conservation laws and the equality of the two forms of the assignment step."""
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(11)
    rows = torch.empty((n, n, 3), dtype=torch.uint8, device="cuda")
    for r0 in range(0, n, 2048):   # smooth ramps + noise: about a million distinct colours
        y = torch.arange(r0, r0 + 2048, device="cuda").view(-1, 1, 1)
        x = torch.arange(n, device="cuda").view(1, -1, 1)
        base = torch.cat([(x * 3 + y) // 197, (x + y * 5) // 311, (x * 7 + y * 2) // 523], dim=2)
        rows[r0:r0 + 2048] = ((base + torch.randint(0, 12, (2048, n, 3), device="cuda", generator=g)) % 256).to(torch.uint8)
    npix = n * n
    init = torch.rand((16, 3), device="cuda", generator=g) * 255
    hist = torch.zeros((1 << 24,), dtype=torch.int32, device="cuda")
    eng.pixels_histogram(rows, hist)
    assert int(hist.to(torch.int64).sum()) == npix
    keys, counts = eng.histogram_compact(hist)
    assert int(counts.to(torch.int64).sum()) == npix and bool((keys[1:] > keys[:-1]).all())
    parts = [eng.histogram_compact(hist, p, 8) for p in range(8)]
    assert sum(int(k.numel()) for k, _ in parts) == keys.numel()
    a = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    lab_px = torch.empty((npix,), dtype=torch.uint8, device="cuda")
    eng.kmeans_pixels_step(rows, init, a, lab_px)
    b = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    for k, c in parts:                     # the eight colour shares add up to the per-pixel sums
        eng.kmeans_hist_step(k, c, init, b)
    assert torch.equal(a, b) and int(a[:, 3].sum()) == npix
    tot = torch.zeros((3,), dtype=torch.int64, device="cuda")
    for r0 in range(0, n, 2048):
        tot += rows[r0:r0 + 2048].reshape(-1, 3).to(torch.int64).sum(0)
    assert torch.equal(a[:, :3].sum(0), tot.flip(0))
    lab_e = torch.empty((keys.numel(),), dtype=torch.uint8, device="cuda")
    c2 = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    eng.kmeans_hist_step(keys, counts, init, c2, lab_e)
    lut = torch.zeros((1 << 24,), dtype=torch.uint8, device="cuda")
    eng.hist_labels_to_lut(keys, lab_e, lut)
    lab2 = torch.empty_like(lab_px)
    eng.pixels_lookup(rows, lut, lab2)
    assert torch.equal(lab_px, lab2)
    assert torch.equal(torch.bincount(lab2.to(torch.int64), minlength=16), a[:, 3])
    # the farthest-member search with the global index base of the last row shard of 8
    far = torch.zeros((1,), dtype=torch.int64, device="cuda")
    r0 = n - n // 8
    donor = int(a[:, 3].argmax())
    base3 = [float(v) for v in (a[donor, :3].to(torch.float64) / float(a[donor, 3])).to(torch.float32)]
    eng.kmeans_pixels_farthest(rows[r0:], init, donor, base3, r0 * n, far)
    code = int(far.item())
    assert code > 0
    gidx = (code - 1) & 0xFFFFFFFF
    assert r0 * n <= gidx < npix and int(lab_px[gidx]) == donor


def test_p2p_update_kernel_on_one_rank_equals_plain_update(eng):
    """llfe_kmeans_update_p2p with a world of one (its own mailbox only): same centres / state / totals as
    llfe_kmeans_update over several epochs (both mailbox parities), and the converged state makes it a no-op."""
    k = 16
    rng = np.random.default_rng(0)
    mailbox = eng.raw_malloc(int(eng.ctx.lib.llfe_p2p_mailbox_bytes()))
    table = torch.tensor([mailbox], dtype=torch.int64, device="cuda")
    c_a = torch.from_numpy(rng.random((k, 3)).astype(np.float32) * 255).cuda()
    c_b = c_a.clone()
    st_a = torch.zeros(4, dtype=torch.int32, device="cuda")
    st_b = st_a.clone()
    sh_a = torch.zeros(1, dtype=torch.float64, device="cuda")
    sh_b = sh_a.clone()
    tot_a = torch.zeros((k, 4), dtype=torch.int64, device="cuda")
    tot_b = tot_a.clone()
    for it in range(5):
        cnt = rng.integers(1, 1 << 28, (k, 1))
        sums = np.concatenate([cnt * rng.integers(0, 256, (k, 3)), cnt], axis=1).astype(np.int64)
        if it >= 3:
            sums = last                                  # repeated sums: zero shift -> converged
        last = sums
        pa = torch.from_numpy(sums).cuda()
        pb = pa.clone()
        eng.kmeans_update(pa, c_a, st_a, sh_a, max_iter=200, eps=0.2, consumed=tot_a, zero_sums=True)
        eng.kmeans_update_p2p(pb, table, 0, 1, c_b, st_b, sh_b, tot_b, max_iter=200, eps=0.2)
        assert torch.equal(c_a, c_b) and torch.equal(st_a, st_b) and torch.equal(tot_a, tot_b) and torch.equal(sh_a, sh_b)
        assert torch.equal(pa, pb)                       # cleared while running, left alone once converged
    assert int(st_b[1]) == 1
    eng.raw_free(mailbox)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box (gpurun --gpus 2)")
def test_p2p_fit_equals_allreduce_fit_on_two_gpus():
    """The fused NVLink exchange gives bit-identical centres to the NCCL all-reduce loop (tools/p2p_check.py)."""
    import json
    import subprocess

    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tools", "p2p_check.py"),
                          "--size", "2048"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["identical"] and rec["world"] == 2
