"""Host logic of services.batching.RequestBatcher with a fake analyzer (no GPU): grouping by shape, batch
size and wait limits, per-request results, error isolation."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from low_level_feature_extraction_b200.services.batching import RequestBatcher  # noqa: E402
from low_level_feature_extraction_b200.services.models import ColorFeatures  # noqa: E402


class FakeAnalyzer:
    """run_host: masks that encode each image's first pixel, two fixed clusters whose sizes depend on the image."""
    calls = []

    def __init__(self, h, w, fail_on=None):
        self.h, self.w, self.fail_on = h, w, fail_on

    def run_host(self, images, host_out=None):
        n = images.shape[0]
        FakeAnalyzer.calls.append(((self.h, self.w), n))
        if self.fail_on is not None and n >= self.fail_on:
            raise RuntimeError("boom")
        tag = images[:, 0, 0, 0].clone()
        mask = tag.view(n, 1, 1).expand(n, self.h, self.w).contiguous()
        centers = torch.zeros((n, 5, 3), dtype=torch.float32)
        centers[:, 0] = torch.tensor([200.9, 10.2, 10.7])
        centers[:, 1] = torch.tensor([10.0, 10.0, 250.0])
        sizes = torch.zeros((n, 5), dtype=torch.int32)
        sizes[:, 0] = 3
        sizes[:, 1] = tag.to(torch.int32)            # image tag > 3: the second cluster becomes primary
        return {"shape_mask": mask, "shadow_mask": 255 - mask, "centers": centers,
                "k_used": torch.full((n,), 2, dtype=torch.int32), "cluster_sizes": sizes,
                "shadow_sums": torch.stack([tag.to(torch.int64) * 100, torch.full((n,), 1, dtype=torch.int64)], dim=1),
                "count": torch.full((n,), 7, dtype=torch.int32)}


def _img(h, w, tag):
    a = np.zeros((h, w, 3), np.uint8)
    a[0, 0, 0] = tag
    return a


def test_groups_by_shape_and_resolves_every_request():
    FakeAnalyzer.calls = []
    with RequestBatcher(max_batch=8, max_wait_ms=200.0, shapes=False, analyzer_factory=FakeAnalyzer) as rb:
        futs = [rb.submit(_img(6, 8, t)) for t in (1, 2, 9)] + [rb.submit(_img(5, 4, t)) for t in (4, 2)]
        res = [f.result(timeout=30) for f in futs]
    for tag, r in zip((1, 2, 9, 4, 2), res):
        assert isinstance(r["colors"], ColorFeatures) and r["shapes"] is None
        assert r["shape_mask"][0, 0] == tag and r["shadow_mask"][0, 0] == 255 - tag
        # palette order follows the cluster sizes: tag > 3 -> the blue cluster leads; centres truncate like astype(uint8)
        assert r["colors"].primary == ("#0a0afa" if tag > 3 else "#c80a0a")
        assert r["shadow_level"] == ("Low" if 255 - tag * 100 < 30 else "Moderate" if 255 - tag * 100 < 60 else "High")
    assert sorted(FakeAnalyzer.calls) == [((5, 4), 2), ((6, 8), 3)]     # one launch per shape
    assert res[0]["shape_mask"].shape == (6, 8) and res[3]["shape_mask"].shape == (5, 4)


def test_max_batch_splits_and_counts():
    FakeAnalyzer.calls = []
    rb = RequestBatcher(max_batch=4, max_wait_ms=100.0, shapes=False, analyzer_factory=FakeAnalyzer)
    futs = [rb.submit(_img(4, 4, t)) for t in range(1, 11)]
    for t, f in zip(range(1, 11), futs):
        assert f.result(timeout=30)["shape_mask"][0, 0] == t
    rb.close()
    rb.close()                                                        # idempotent
    assert all(n <= 4 for _, n in FakeAnalyzer.calls) and sum(n for _, n in FakeAnalyzer.calls) == 10
    assert rb.images == 10 and rb.batches == len(FakeAnalyzer.calls)
    with pytest.raises(RuntimeError):
        rb.submit(_img(4, 4, 1))


def test_errors_stay_with_their_requests():
    with RequestBatcher(max_batch=8, max_wait_ms=100.0, shapes=False,
                        analyzer_factory=lambda h, w: FakeAnalyzer(h, w, fail_on=2 if h == 3 else None)) as rb:
        bad_input = rb.submit(np.zeros((4, 4), np.uint8))             # not BGR: fails alone, immediately
        ok = rb.submit(_img(4, 4, 5))
        boom = [rb.submit(_img(3, 3, 1)), rb.submit(_img(3, 3, 2))]   # this shape's launch raises
        assert ok.result(timeout=30)["shape_mask"][0, 0] == 5
        with pytest.raises(__import__("cv2").error):               # what the reference's cvtColor raises for it
            bad_input.result(timeout=30)
        for f in boom:
            with pytest.raises(RuntimeError, match="boom"):
                f.result(timeout=30)


def test_concurrent_submitters_share_launches():
    FakeAnalyzer.calls = []
    with RequestBatcher(max_batch=16, max_wait_ms=300.0, shapes=False, analyzer_factory=FakeAnalyzer) as rb:
        out = {}
        start = threading.Barrier(8)

        def client(t):
            start.wait()
            out[t] = rb.analyze(_img(4, 6, t))

        th = [threading.Thread(target=client, args=(t,)) for t in range(1, 9)]
        for x in th:
            x.start()
        for x in th:
            x.join()
    assert all(out[t]["shape_mask"][0, 0] == t for t in range(1, 9))
    assert len(FakeAnalyzer.calls) < 8                                 # requests were grouped
