"""Multi-rank host logic on CPU: world_size-2 gloo process groups (no GPU needed).

The arithmetic backend here is the oracle stand-in of tests/cpu_backend.py; what is under
test is low_level_feature_extraction_b200.dist: sharding, the per-iteration all-reduce of
the K x 4 accumulator, identical decisions on every rank, and the cross-shard repair.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from low_level_feature_extraction_b200 import dist as ldist  # noqa: E402
from low_level_feature_extraction_b200.synth import design_image  # noqa: E402
from oracle import cvops  # noqa: E402


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 256, 8192, 1080):
        for ws in (1, 2, 3, 4, 8):
            parts = [ldist.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ldist.shard_range(4, 2, 2)


def _case(kind):
    if kind == "design":
        img = design_image(48, 64, 5)
        init = np.float32(img.reshape(-1, 3)[:, ::-1][np.random.default_rng(42).choice(48 * 64, 6, replace=False)])
    else:  # two identical initial centres -> an empty cluster on the first update -> repair across shards
        img = design_image(40, 56, 9)
        px = img.reshape(-1, 3)[:, ::-1]
        init = np.float32(px[np.random.default_rng(7).choice(len(px), 5, replace=False)])
        init[3] = init[1]
        init[4] = init[1]
    return img, init


def _worker(rank, world, port, kind, histogram, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_backend import OracleBackend

        img, init = _case(kind)
        h, w, _ = img.shape
        r0, r1 = ldist.row_shard(h, rank, world)
        rows = torch.from_numpy(np.ascontiguousarray(img[r0:r1]))
        km = ldist.PixelKMeans(OracleBackend(), histogram=histogram)
        res = km.fit(rows, torch.from_numpy(init), index_base=r0 * w, want_labels=True)
        lab = ldist.gather_results(res.labels.reshape(r1 - r0, w), h)
        q.put((rank, res.centers.numpy().copy(), res.iters, res.sums_counts.numpy().copy(), lab.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("histogram", [True, False])
@pytest.mark.parametrize("kind", ["design", "empty_cluster"])
@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_kmeans_equals_single_process_oracle(kind, world, histogram):
    img, init = _case(kind)
    px = img.reshape(-1, 3)[:, ::-1]
    c_ref, l_ref, it_ref, s_ref, n_ref = cvops.lloyd_exact(px, init)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, histogram, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, centers, iters, sums, labels in got:
        assert iters == it_ref
        assert np.array_equal(centers, c_ref)                      # bit-identical on every rank
        assert np.array_equal(sums[:, :3], s_ref) and np.array_equal(sums[:, 3], n_ref)
        assert np.array_equal(labels.reshape(-1), l_ref.astype(np.uint8))


def test_single_process_is_identity_collective():
    from cpu_backend import OracleBackend

    img, init = _case("design")
    px = img.reshape(-1, 3)[:, ::-1]
    c_ref, l_ref, it_ref, _, _ = cvops.lloyd_exact(px, init)
    for histogram in (True, False):
        res = ldist.PixelKMeans(OracleBackend(), histogram=histogram).fit(torch.from_numpy(img), torch.from_numpy(init),
                                                                          want_labels=True)
        assert res.iters == it_ref and np.array_equal(res.centers.numpy(), c_ref)
        assert np.array_equal(res.labels.numpy(), l_ref.astype(np.uint8))


def test_host_pipeline_stage_schedule():
    """BatchAnalyzer.run_host: stages cover the batch exactly once, never exceed host_chunk, ramp up and down."""
    from low_level_feature_extraction_b200.batch import BatchAnalyzer, BatchConfig

    class Host:   # _stages only needs cfg
        pass

    for chunk in (1, 4, 16, 32):
        h = Host()
        h.cfg = BatchConfig(host_chunk=chunk)
        for n in (1, 2, 3, 15, 16, 17, 40, 100, 256, 1000):
            st = BatchAnalyzer._stages(h, n)
            assert [i0 for i0, _ in st] == [sum(m for _, m in st[:j]) for j in range(len(st))]
            assert sum(m for _, m in st) == n and all(0 < m <= chunk for _, m in st)
            if n >= 8 * chunk and chunk >= 4:
                assert st[0][1] < chunk and st[-1][1] < chunk and max(m for _, m in st) == chunk
