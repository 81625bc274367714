"""Oracle-backed stand-in for the three per-pixel k-means entry points of `ops.Engine`
(`kmeans_pixels_step`, `kmeans_pixels_farthest`, `kmeans_update`) on CPU tensors.

TEST INFRASTRUCTURE ONLY: it lets the gloo (world_size 2) tests run the host logic of
`low_level_feature_extraction_b200.dist.PixelKMeans` -- iteration bookkeeping, the
all-reduce, the cross-shard empty-cluster repair -- in a container without a GPU.  The
product backend is always the CUDA engine; nothing under the package imports this file.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import cvops

f32 = np.float32


class OracleBackend:
    def kmeans_pixels_zero(self, sums, state=None):
        if state is not None and (int(state[1]) or int(state[3])):
            return
        sums.zero_()

    def kmeans_pixels_step(self, bgr_rows, centers, sums, labels=None, state=None):
        if state is not None and (int(state[1]) or int(state[3])):
            return
        px = bgr_rows.reshape(-1, 3).numpy()[:, ::-1]  # RGB
        if len(px) == 0:
            return
        lab, _ = cvops.assign(px.astype(f32), centers.numpy())
        k = centers.shape[0]
        s = sums.numpy()
        for j in range(3):
            s[:, j] += np.bincount(lab, weights=px[:, j].astype(np.float64), minlength=k).astype(np.int64)
        s[:, 3] += np.bincount(lab, minlength=k)
        if labels is not None:
            labels.numpy()[:] = lab.astype(np.uint8)

    def kmeans_pixels_farthest(self, flat_bgr, centers, donor, base3, index_base, out, skip=()):
        px = flat_bgr.reshape(-1, 3).numpy()[:, ::-1]
        if len(px) == 0:
            return
        lab, _ = cvops.assign(px.astype(f32), centers.numpy())
        members = np.flatnonzero(lab == donor)
        members = members[~np.isin(members + index_base, np.asarray(list(skip), dtype=np.int64))]
        if len(members) == 0:
            return
        d = cvops.l2sqr(px[members].astype(f32), np.asarray(base3, f32))
        code = (d.view(np.uint32).astype(np.int64) << 32 | (members + index_base)) + 1
        out[0] = max(int(out[0]), int(code.max()))

    def kmeans_update(self, sums, centers, state, shift, max_iter=200, eps=0.2):
        if int(state[1]) or int(state[3]):
            return
        s = sums.numpy()
        n_empty = int((s[:, 3] == 0).sum())
        state[2] = n_empty
        if n_empty:
            state[3] = 1
            return
        new = (s[:, :3].astype(np.float64) / s[:, 3:4].astype(np.float64)).astype(f32)
        sh = cvops.center_shift(new, centers.numpy())
        centers.copy_(torch.from_numpy(new))
        it0 = int(state[0])
        state[0] = it0 + 1
        state[1] = int(it0 + 1 == max(max_iter, 2) or (it0 > 0 and sh <= eps * eps))
        shift[0] = sh
