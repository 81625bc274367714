"""Oracle-backed stand-in for the per-pixel k-means entry points of `ops.Engine`
(`kmeans_pixels_step`, `kmeans_pixels_farthest`, `kmeans_update`, and the colour-histogram
form `pixels_histogram` ... `pixels_lookup`) on CPU tensors.

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

    def kmeans_pixels_farthest(self, flat_bgr, centers, donor, base3, index_base, out, skip=(), want_dist_bits=0):
        px = flat_bgr.reshape(-1, 3).numpy()[:, ::-1]
        if len(px) == 0:
            return
        lab, _ = cvops.assign(px.astype(f32), centers.numpy())
        members = np.flatnonzero(lab == donor)
        members = members[~np.isin(members + index_base, np.asarray(list(skip), dtype=np.int64))]
        if len(members) == 0:
            return
        d = cvops.l2sqr(px[members].astype(f32), np.asarray(base3, f32))
        keep = (d.view(np.uint32) == np.uint32(want_dist_bits)) if want_dist_bits else np.ones(len(d), bool)
        members, d = members[keep], d[keep]
        if len(members) == 0:
            return
        code = (d.view(np.uint32).astype(np.int64) << 32 | (members + index_base)) + 1
        out[0] = max(int(out[0]), int(code.max()))

    def kmeans_hist_farthest(self, keys, centers, donor, base3, out_bits):
        if keys.numel() == 0:
            return
        kk = keys.numpy().astype(np.int64)
        px = np.stack([kk >> 16, (kk >> 8) & 255, kk & 255], axis=1).astype(f32)
        lab, _ = cvops.assign(px, centers.numpy())
        members = np.flatnonzero(lab == donor)
        if len(members) == 0:
            return
        d = cvops.l2sqr(px[members], np.asarray(base3, f32))
        out_bits[0] = max(int(out_bits[0]), int(d.view(np.uint32).max()))

    def kmeans_update(self, sums, centers, state, shift, max_iter=200, eps=0.2, consumed=None, zero_sums=False):
        if int(state[1]) or int(state[3]):
            return
        if consumed is not None:
            consumed.copy_(sums)
            if zero_sums:
                sums.zero_()
            sums = consumed
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

    # ---- colour-histogram form ---------------------------------------------------------------------
    HB = 2048   # keys per ownership block, as in k_colorhist.cu

    def pixels_histogram(self, bgr_rows, hist):
        px = bgr_rows.reshape(-1, 3).numpy().astype(np.int64)
        if len(px) == 0:
            return
        key = (px[:, 2] << 16) | (px[:, 1] << 8) | px[:, 0]
        hist.numpy()[:] += np.bincount(key, minlength=1 << 24).astype(np.int32)

    def histogram_compact(self, hist, part=0, parts=1):
        h = hist.numpy()
        keys = np.flatnonzero(h)
        keys = keys[(keys // self.HB) % parts == part]
        return torch.from_numpy(keys.astype(np.int32)), torch.from_numpy(h[keys].astype(np.int32))

    def histogram_compact_device(self, hist, part=0, parts=1, packed=False, cap=None):
        """(keys, counts) padded to the capacity of the part + the entry count as a tensor, as the CUDA engine returns."""
        h = hist.numpy()
        if packed:      # slot j of the share is block j * parts + part of the table
            nz = np.flatnonzero(h)
            keys = ((nz // self.HB) * parts + part) * self.HB + nz % self.HB
            vals = h[nz]
        else:
            keys = np.flatnonzero(h)
            keys = keys[(keys // self.HB) % parts == part]
            vals = h[keys]
        pad = 5    # some unused capacity behind the entries, like the real buffers
        k = np.concatenate([keys, np.zeros(pad, np.int64)]).astype(np.int32)
        v = np.concatenate([vals, np.zeros(pad, np.int64)]).astype(np.int32)
        return torch.from_numpy(k), torch.from_numpy(v), torch.tensor([len(keys)], dtype=torch.int32)

    def kmeans_hist_step(self, keys, counts, centers, sums, labels=None, state=None, n_dev=None):
        if state is not None and (int(state[1]) or int(state[3])):
            return
        if n_dev is not None:
            nn = int(n_dev[0])
            keys, counts = keys[:nn], counts[:nn]
            labels = labels[:nn] if labels is not None else None
        if keys.numel() == 0:
            return
        kk = keys.numpy().astype(np.int64)
        px = np.stack([kk >> 16, (kk >> 8) & 255, kk & 255], axis=1)   # RGB
        lab, _ = cvops.assign(px.astype(f32), centers.numpy())
        k = centers.shape[0]
        s = sums.numpy()
        cnt = counts.numpy().astype(np.int64)
        for j in range(3):
            np.add.at(s[:, j], lab, cnt * px[:, j])
        np.add.at(s[:, 3], lab, cnt)
        if labels is not None:
            labels.numpy()[:] = lab.astype(np.uint8)

    def hist_labels_to_lut(self, keys, labels, lut):
        lut.numpy()[keys.numpy()] = labels.numpy()

    def pixels_lookup(self, bgr_rows, lut, labels):
        px = bgr_rows.reshape(-1, 3).numpy().astype(np.int64)
        labels.numpy()[:] = lut.numpy()[(px[:, 2] << 16) | (px[:, 1] << 8) | px[:, 0]]
