"""Multi-GPU plumbing: one process per GPU, torch.distributed for the (few) exchanges.

Two ways the path shards:

* batches of images (BASELINE configs 2-4): images are independent, rank r takes
  `shard_range(n, r, world)`; there is NO data-path collective, only an optional
  gather of the tiny per-image results (`gather_results`);
* one oversized image (config 5): pixel rows are split across ranks
  (`row_shard`), every Lloyd iteration each rank accumulates exact uint64
  {sum R, sum G, sum B, count} per cluster over its rows on the device
  (`llfe_kmeans_pixels_step`), the K x 4 accumulator is all-reduced
  (ncclSum over NVLink: 512 B at K = 16), and every rank then runs the identical
  centre update + convergence test (`llfe_kmeans_update`) -- integer sums make the
  result independent of the number of ranks, bit for bit.

`PixelKMeans` is written against a small backend interface (`step`, `farthest`,
`update`) so that the host logic -- iteration bookkeeping, the all-reduce, cv2's
empty-cluster repair across shards -- is exercised on CPU with gloo in tests/ while
the product backend is always `ops.Engine` (CUDA, no fallback).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced split of n items: the first n % world ranks get one extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world: {rank}/{world_size}")
    base, extra = divmod(n, world_size)
    i0 = rank * base + min(rank, extra)
    return i0, i0 + base + (1 if rank < extra else 0)


def row_shard(h: int, rank: int, world_size: int) -> tuple[int, int]:
    """Row range [r0, r1) of an h-row image owned by `rank`."""
    return shard_range(h, rank, world_size)


def gather_results(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-image results (palettes, shadow sums, ...) of a batch sharded with
    `shard_range` on every rank.  Not on the hot path: a few bytes per image."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    cap = max(b - a for a, b in sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


class P2PSums:
    """Peer mailboxes for the fused all-reduce + centre update of config 5 (`llfe_kmeans_update_p2p`): every rank of a
    one-node NCCL group allocates a mailbox in its own HBM, exports it as a CUDA IPC handle, and maps the others'.
    Built once per (engine, group); `table` is the device array of the `world` mailbox pointers."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("at most 16 ranks")
        # one exported allocation per rank: the mailbox, then this rank's 2^24-bin colour table (64 MiB)
        self.table_offset = (int(engine.ctx.lib.llfe_p2p_mailbox_bytes()) + 255) & ~255
        nbytes = self.table_offset + (1 << 24) * 4
        self.own = engine.raw_malloc(nbytes)
        torch.cuda.synchronize(engine.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, engine.ipc_export(self.own), group=group)
        self.peers = [self.own if r == self.rank else engine.ipc_open(handles[r]) for r in range(self.world)]
        self.table = torch.tensor(self.peers, dtype=torch.int64, device=engine.device)
        self.hist_tables = torch.tensor([p + self.table_offset for p in self.peers], dtype=torch.int64, device=engine.device)
        self.own_hist = self.own + self.table_offset
        dist.barrier(group=group)      # nobody writes into a mailbox before its owner has zeroed it

    def close(self):
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.engine.ipc_close(p)
        self.engine.raw_free(self.own)
        self.peers = []


@dataclass
class PixelKMeansResult:
    centers: torch.Tensor      # (k, 3) float32, RGB
    iters: int
    sums_counts: torch.Tensor  # (k, 4) int64: sum R, sum G, sum B, count (global)
    shift: float
    labels: torch.Tensor | None = None  # (rows*w,) uint8 for this rank's rows, if requested


class PixelKMeans:
    """Per-pixel Lloyd k-means of ONE image whose rows are sharded across ranks.

    The arithmetic is the pinned exact-sum rule of SURVEY.md A.8: float32 assignment
    (separate mul/add, strict <), integer sums, c = float32(double(sum)/double(count)),
    shift test max_k |c - old|^2 <= eps^2 from iteration 1 on, cv2's empty-cluster repair.
    """

    def __init__(self, backend, group=None, iterations_per_sync: int = 8, histogram: bool = True, p2p: bool | None = None):
        self.be = backend
        self.group = group
        # p2p: the per-iteration all-reduce of the k x 4 sums + the centre update as ONE kernel over NVLink peer stores
        # (P2PSums) instead of ncclAllReduce + update.  None = whenever the group is NCCL with more than one rank and the
        # backend is the CUDA engine; the gloo / CPU-backend tests keep the all-reduce.
        self.p2p = p2p
        self._p2p_sums = None
        self.persistent = None    # None / True: the histogram form iterates inside one persistent kernel when it can
        self.iterations_per_sync = iterations_per_sync   # iterations enqueued per host look at the device state
        # histogram=True: stream the rows once into a 2^24-bin colour count table, all-reduce it, and iterate
        # over this rank's share of the DISTINCT colours weighted by their counts (identical labels, sums and
        # centres; the image is not re-read every iteration).  False: every iteration re-reads the rank's rows.
        self.histogram = histogram

    def _rank_world(self) -> tuple[int, int]:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    # -- collectives (identity in a single process) ----------------------------------------
    def _allreduce(self, t: torch.Tensor, op) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=op, group=self.group)

    def _p2p_setup(self):
        if self.p2p is False or self._p2p_sums is not None:
            return self._p2p_sums
        ok = (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
              and dist.get_backend(self.group) == "nccl" and hasattr(self.be, "kmeans_update_p2p"))
        if not ok:
            if self.p2p:
                raise RuntimeError("p2p=True needs an NCCL group with more than one rank and the CUDA engine")
            return None
        self._p2p_sums = P2PSums(self.be, self.group)
        return self._p2p_sums

    HB = 2048                    # keys per ownership block of the colour table (k_colorhist.cu)

    def _reduce_scatter_blocks(self, hist: torch.Tensor, rank: int, ws: int):
        """Sum of the colour table over the ranks, but only this rank's blocks (b % ws == rank), packed back to back.
        None when the world size does not divide the block count or the backend has no reduce-scatter (gloo): the
        caller falls back to the all-reduce of the whole table."""
        nblk = hist.numel() // self.HB
        if ws == 1:
            return hist                          # packed with parts = 1 is the table itself
        if nblk % ws != 0 or not (dist.is_available() and dist.is_initialized()):
            return None
        if dist.get_backend(self.group) != "nccl":
            return None
        send = hist.view(nblk // ws, ws, self.HB).transpose(0, 1).contiguous().view(-1)   # rank-major blocks
        share = torch.empty((hist.numel() // ws,), dtype=hist.dtype, device=hist.device)
        dist.reduce_scatter_tensor(share, send, op=dist.ReduceOp.SUM, group=self.group)
        return share

    def fit(self, bgr_rows: torch.Tensor, init_centers: torch.Tensor, index_base: int = 0, max_iter: int = 200,
            eps: float = 0.2, want_labels: bool = False) -> PixelKMeansResult:
        """bgr_rows: this rank's (rows, w, 3) uint8 BGR rows (may be empty); init_centers: (k, 3)
        float32 RGB, identical on every rank; index_base: global pixel index of the first local pixel."""
        be = self.be
        dev = bgr_rows.device
        k = int(init_centers.shape[0])
        centers = init_centers.to(device=dev, dtype=torch.float32).contiguous().clone()
        local = torch.zeros((k, 4), dtype=torch.int64, device=dev)    # this rank's partial sums
        sums = torch.zeros((k, 4), dtype=torch.int64, device=dev)     # all-reduced sums
        state = torch.zeros((4,), dtype=torch.int32, device=dev)      # iteration, converged, n_empty, frozen
        shift = torch.zeros((1,), dtype=torch.float64, device=dev)
        far = torch.zeros((1,), dtype=torch.int64, device=dev)
        npix = bgr_rows.numel() // 3
        labels = torch.empty((npix,), dtype=torch.uint8, device=dev) if want_labels else None
        flat = bgr_rows.reshape(-1, 3)
        if self.histogram:
            rank, ws = self._rank_world()
            p2p_tab = self._p2p_setup() if (1 << 24) // self.HB % max(ws, 1) == 0 else None
            if p2p_tab is not None:
                # the table lives in memory the peers can read: device barrier, then every rank pulls and sums ITS share
                # (the interleaved 2048-key blocks b % G == rank) straight out of the peers' HBM over NVLink
                be.raw_memset(p2p_tab.own_hist, 0, (1 << 24) * 4)
                be.pixels_histogram_raw(bgr_rows, p2p_tab.own_hist)
                be.p2p_barrier(p2p_tab.table, rank, ws)
                hist = None
                share = torch.empty(((1 << 24) // ws,), dtype=torch.int32, device=dev)
                be.histogram_pull_reduce(p2p_tab.hist_tables, rank, ws, share)
            else:
                hist = torch.zeros((1 << 24,), dtype=torch.int32, device=dev)
                be.pixels_histogram(bgr_rows, hist)
                # Every rank needs only ITS share of the summed table: reduce-scatter over the block-transposed table
                # moves 1/G of what an all-reduce of the 64 MiB would.
                share = self._reduce_scatter_blocks(hist, rank, ws)
            if share is not None:
                keys, counts, n_dev = be.histogram_compact_device(share, rank, ws, packed=True)
            else:
                self._allreduce(hist, dist.ReduceOp.SUM)
                keys, counts, n_dev = be.histogram_compact_device(hist, rank, ws)
            del hist, share
            # the entry count stays on the device: no host round trip between the compaction and the iterations
            entry_labels = torch.empty((keys.numel(),), dtype=torch.uint8, device=dev) if want_labels else None

            def step():
                be.kmeans_hist_step(keys, counts, centers, local, entry_labels, state, n_dev)
        else:
            def step():
                be.kmeans_pixels_step(bgr_rows, centers, local, labels, state)

        p2p = self._p2p_setup()
        _, ws_now = self._rank_world()
        persistent = (self.histogram and self.persistent is not False and hasattr(be, "kmeans_hist_lloyd")
                      and (ws_now == 1 or p2p is not None))
        overrides: list[tuple[int, int]] = []   # (local pixel, label) set by the repair after the last assignment
        overrides_iter = -1
        batch = max(1, int(self.iterations_per_sync))
        while True:
            # `batch` iterations are enqueued back to back: assignment into the per-rank accumulator, in-place
            # all-reduce, update (which keeps the totals in `sums` and clears the accumulator).  Once the device
            # state says converged (or frozen for a repair) the remaining ones are no-ops on every rank alike.
            if persistent:
                # the whole batch is one persistent cooperative kernel (assignment, NVLink exchange, update per iteration)
                be.kmeans_hist_lloyd(keys, counts, centers, local, entry_labels, state, shift, sums, n_dev=n_dev,
                                     mailboxes=None if p2p is None else p2p.table, rank=0 if p2p is None else p2p.rank,
                                     world=1 if p2p is None else p2p.world, max_iter=max_iter, eps=eps,
                                     iterations=4 * batch)
            for _ in range(0 if persistent else batch):
                step()
                if p2p is not None:
                    be.kmeans_update_p2p(local, p2p.table, p2p.rank, p2p.world, centers, state, shift, sums,
                                         max_iter=max_iter, eps=eps)
                else:
                    self._allreduce(local, dist.ReduceOp.SUM)
                    be.kmeans_update(local, centers, state, shift, max_iter=max_iter, eps=eps, consumed=sums,
                                     zero_sums=True)
            st = state.tolist()          # the one host synchronisation per batch
            if st[3]:
                if self.histogram and keys.numel() != int(n_dev.item()):   # the repair walks the real list (rare path)
                    nn = int(n_dev.item())
                    keys, counts = keys[:nn], counts[:nn]
                    entry_labels = entry_labels[:nn] if entry_labels is not None else None
                overrides = self._repair(flat, centers, sums, far, index_base, npix, None if self.histogram else labels,
                                         keys if self.histogram else None)
                state[2:4] = 0
                be.kmeans_update(sums, centers, state, shift, max_iter=max_iter, eps=eps)
                st = state.tolist()
                overrides_iter = st[0]
                if st[3]:
                    raise RuntimeError("k-means: an empty cluster survived the repair (fewer distinct pixels than k?)")
            if st[1]:
                if self.histogram and want_labels:
                    lut = torch.zeros((1 << 24,), dtype=torch.uint8, device=dev)
                    nn = int(n_dev.item())
                    be.hist_labels_to_lut(keys[:nn], entry_labels[:nn], lut)
                    self._allreduce(lut, dist.ReduceOp.SUM)   # every colour belongs to exactly one rank's share
                    be.pixels_lookup(bgr_rows, lut, labels)
                    if overrides_iter == st[0]:               # the repair came after the last assignment
                        for li, j in overrides:
                            labels[li] = j
                return PixelKMeansResult(centers, st[0], sums, float(shift.item()), labels)

    def _repair(self, flat, centers, sums, far, index_base, npix, labels, keys=None) -> list[tuple[int, int]]:
        """cv2's empty-cluster repair on the all-reduced sums: for each empty cluster (in order) the
        biggest cluster (first max) gives up its member farthest from its provisional mean (last max
        wins => highest global pixel index).  `centers` still holds the centres the labels came from."""
        k = centers.shape[0]
        moved: list[int] = []
        overrides: list[tuple[int, int]] = []
        host = sums.cpu()
        assigned_from = centers.clone()  # `centers` gets the donors' provisional means below
        for j in range(k):
            if int(host[j, 3]) != 0:
                continue
            cnt = host[:, 3]
            donor = 0
            for k1 in range(1, k):
                if int(cnt[donor]) < int(cnt[k1]):
                    donor = k1
            base = (host[donor, :3].to(torch.float64) / float(cnt[donor])).to(torch.float32)
            base3 = [float(v) for v in base]
            dmin = 0
            if keys is not None:
                # the farthest COLOUR of the donor comes from the distinct-colour list; the pass over the rows then
                # only has to find the last pixel at exactly that distance from `base`
                bits = torch.zeros((1,), dtype=torch.int32, device=flat.device)
                self.be.kmeans_hist_farthest(keys, assigned_from, donor, base3, bits)
                self._allreduce(bits, dist.ReduceOp.MAX)
                dmin = int(bits.item())
            code = 0
            for threshold in ((dmin, 0) if dmin else (0,)):   # 0: every pixel of that colour was already moved
                far.zero_()
                self.be.kmeans_pixels_farthest(flat, assigned_from, donor, base3, index_base, far, skip=moved,
                                               want_dist_bits=threshold)
                self._allreduce(far, dist.ReduceOp.MAX)
                code = int(far.item())
                if code:
                    break
            if code == 0:
                raise RuntimeError("k-means repair: the donor cluster has no member")
            gidx = (code - 1) & 0xFFFFFFFF
            # the owner contributes the pixel's colour (RGB); everyone else zeros
            px = torch.zeros((3,), dtype=torch.int64, device=flat.device)
            li = gidx - index_base
            if 0 <= li < npix:
                px = flat[li].flip(0).to(torch.int64)
                overrides.append((li, j))
                if labels is not None:
                    labels[li] = j
            self._allreduce(px, dist.ReduceOp.SUM)
            pxh = px.cpu()
            host[donor, :3] -= pxh
            host[donor, 3] -= 1
            host[j, :3] += pxh
            host[j, 3] += 1
            moved.append(gidx)
            # OpenCV stores the donor's provisional mean in old_centers[donor]: the shift is measured from it
            centers[donor] = base.to(centers.device)
        sums.copy_(host)
        return overrides
