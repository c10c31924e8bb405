"""Multi-GPU plumbing: one process per GPU, independent work units, one gather.

The hot path shards without any data-path collective: chromosome arms are independent in
segmentation (cnvlib/segmentation/__init__.py:158-179 maps them over a process pool; every R call
even re-seeds, cbs.py:49) and samples are independent in fix + segment (cnvlib/commands.py:375-403).
So a rank gets a subset of arms (one WGS sample) or of samples (a cohort), runs the ordinary
single-GPU calls on them, and the per-rank segment tables are gathered at the end --
``all_gather`` of the row counts, then ``gather`` of zero-padded payloads.  Works with the NCCL
backend (CUDA tensors) and with gloo (CPU tensors; used by the world_size-2 tests).
"""
from __future__ import annotations

import numpy as np


def lpt_assign(costs, n_shards: int) -> np.ndarray:
    """Longest-processing-time assignment: heaviest unit first, always to the lightest shard.
    Deterministic (stable order, lowest shard index on ties) so every rank computes the same map."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(max(int(n_shards), 1))
    owner = np.zeros(len(costs), dtype=np.int64)
    for u in order:
        r = int(np.argmin(load))
        owner[u] = r
        load[r] += costs[u]
    return owner


def arm_owner(arm_offsets, n_shards: int) -> np.ndarray:
    """Arms -> ranks by LPT on n^2 (the first-level arc count of an arm of n bins)."""
    sizes = np.diff(np.asarray(arm_offsets, dtype=np.int64)).astype(np.float64)
    return lpt_assign(sizes * sizes, n_shards)


def block_owner(n_units: int, n_shards: int) -> np.ndarray:
    """Units -> ranks in contiguous blocks of near-equal size (cohort samples)."""
    return (np.arange(n_units, dtype=np.int64) * n_shards) // max(n_units, 1)


def take_arms(log2, weight, arm_offsets, owner, rank: int):
    """This rank's arms, concatenated: (log2, weight, local arm_offsets, global arm ids)."""
    off = np.asarray(arm_offsets, dtype=np.int64)
    mine = np.flatnonzero(np.asarray(owner) == rank)
    sizes = (off[mine + 1] - off[mine]).astype(np.int64)
    if len(mine):
        idx = np.concatenate([np.arange(off[a], off[a + 1]) for a in mine]) if sizes.sum() else np.zeros(0, np.int64)
    else:
        idx = np.zeros(0, np.int64)
    x = np.ascontiguousarray(np.asarray(log2)[idx], dtype=np.float64)
    w = None if weight is None else np.ascontiguousarray(np.asarray(weight)[idx], dtype=np.float64)
    local_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return x, w, local_off, mine


def _world(group=None):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_tables(columns, dst: int = 0, group=None, device=None, capacity=None):
    """Gather per-rank row tables on ``dst``.

    ``columns`` is a list of equal-length 1-d numpy arrays (int or float; all are shipped as one
    float64 [rows, ncols] payload -- bin indices are < 2^53).  Returns, on ``dst``, the list of
    concatenated columns in rank order (dtypes restored); ``None`` on the other ranks.

    With ``capacity`` (an upper bound on any rank's row count, known to every rank) the exchange is a
    single ``gather`` of [capacity + 1, ncols] payloads whose first row carries the row count; without
    it the counts are all-gathered first and the payload is padded to the largest one.
    """
    import torch
    import torch.distributed as dist

    world, rank = _world(group)
    dtypes = [np.asarray(c).dtype for c in columns]
    if world == 1:
        return [np.asarray(c).copy() for c in columns]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    rows = len(columns[0]) if columns else 0
    if capacity is not None:
        if rows > capacity:
            raise ValueError(f"gather_tables: {rows} rows exceed the agreed capacity {capacity}")
        pay = torch.zeros((int(capacity) + 1, max(len(columns), 1)), dtype=torch.float64, device=device)
        pay[0, 0] = float(rows)
        if rows:
            host = np.stack([np.asarray(c, dtype=np.float64) for c in columns], axis=1)
            pay[1:rows + 1, :len(columns)] = torch.from_numpy(host).to(device)
        out = [torch.zeros_like(pay) for _ in range(world)] if rank == dst else None
        dist.gather(pay, out, dst=dst, group=group)
        if rank != dst:
            return None
        hosts = [o.cpu().numpy() for o in out]
        parts = [h[1:int(h[0, 0]) + 1, :len(columns)] for h in hosts]
        full = np.concatenate(parts, axis=0) if parts else np.zeros((0, len(columns)))
        return [full[:, k].astype(dt) for k, dt in enumerate(dtypes)]
    cnt = torch.tensor([rows], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    counts = [int(c.item()) for c in cnts]
    width = max(max(counts), 1)
    pay = torch.zeros((width, len(columns)), dtype=torch.float64, device=device)
    if rows:
        host = np.stack([np.asarray(c, dtype=np.float64) for c in columns], axis=1)
        pay[:rows] = torch.from_numpy(host).to(device)
    out = [torch.zeros_like(pay) for _ in range(world)] if rank == dst else None
    dist.gather(pay, out, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [o[:c].cpu().numpy() for o, c in zip(out, counts)]
    full = np.concatenate(parts, axis=0) if parts else np.zeros((0, len(columns)))
    return [full[:, k].astype(dt) for k, dt in enumerate(dtypes)]


class TableGatherer:
    """Gather of the per-rank segment tables of a BATCH of work units in one collective.

    A rank appends the small tables it produces (one per sample / step) into a preallocated pinned host
    payload -- no device work, no synchronisation -- and `flush()` ships the whole batch: one host->device
    copy, one `all_gather_into_tensor` of fixed-size payloads (row count in the first row), one device->host
    copy on every rank, decoded on `dst`.  This keeps the collective out of the per-unit critical path
    (a 3 ms segmentation step does not wait for a 100 us gather and its launch latency).

    capacity = upper bound on the rows one rank appends between two flushes, agreed by all ranks.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors; world_size-2 tests)."""

    def __init__(self, ncols: int, capacity: int, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.world, self.rank = _world(group)
        self.group, self.ncols, self.capacity = group, int(ncols), int(capacity)
        if device is None:
            on_gpu = self.world > 1 and dist.get_backend(group) == "nccl"
            device = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        self.device = torch.device(device)
        shape = (self.capacity + 1, self.ncols)
        pin = self.device.type == "cuda"
        self.host = torch.zeros(shape, dtype=torch.float64, pin_memory=pin)
        self.view = self.host.numpy()
        self.rows = 0
        if self.world > 1:
            self.dev_in = torch.zeros(shape, dtype=torch.float64, device=self.device)
            self.dev_out = torch.zeros((self.world,) + shape, dtype=torch.float64, device=self.device)
            self.host_out = torch.zeros((self.world,) + shape, dtype=torch.float64, pin_memory=pin)

    def append(self, columns):
        """columns: equal-length 1-d arrays, one per table column (shipped as float64; bin indices < 2^53)."""
        k = len(columns[0]) if len(columns) else 0
        if self.rows + k > self.capacity:
            raise ValueError(f"TableGatherer: {self.rows + k} rows exceed the agreed capacity {self.capacity}")
        for c, col in enumerate(columns):
            self.view[1 + self.rows:1 + self.rows + k, c] = col
        self.rows += k

    def flush(self, dst: int = 0):
        """Returns on `dst` the [rows, ncols] float64 table of all ranks in rank order (None elsewhere)."""
        import torch.distributed as dist

        self.view[0, 0] = float(self.rows)
        rows, self.rows = self.rows, 0
        if self.world == 1:
            return self.view[1:rows + 1].copy()
        self.dev_in.copy_(self.host, non_blocking=True)
        dist.all_gather_into_tensor(self.dev_out.view(-1), self.dev_in.view(-1), group=self.group)
        if self.rank != dst:
            return None
        self.host_out.copy_(self.dev_out)          # one blocking device->host copy on dst
        out = self.host_out.numpy()
        parts = [out[r, 1:int(out[r, 0, 0]) + 1] for r in range(self.world)]
        return np.concatenate(parts, axis=0)


def segment_arms_sharded(log2, weight, arm_offsets, segment_fn, dst: int = 0, group=None, device=None):
    """Segment one sample's arms across the ranks of ``group``.

    ``segment_fn(log2, weight, local_arm_offsets) -> (seg_arm, seg_lo, seg_hi, seg_mean, ...)`` is the
    single-GPU call (``cnvkit_b200.segmentation.cbs.segment_arms`` / ``haar.segment_arms``) with
    half-open *local* bin ranges.  Returns on ``dst`` the table of the whole sample in arm order with
    global arm ids and global bin ranges -- identical to the unsharded call; ``None`` elsewhere.
    """
    world, rank = _world(group)
    off = np.asarray(arm_offsets, dtype=np.int64)
    owner = arm_owner(off, world)
    x, w, local_off, mine = take_arms(log2, weight, off, owner, rank)
    res = segment_fn(x, w, local_off)
    s_arm, s_lo, s_hi, s_mean = (np.asarray(r) for r in res[:4])
    # local -> global: arm ids through `mine`, bin ranges shifted by the arm's global start
    g_arm = mine[s_arm] if len(s_arm) else np.zeros(0, np.int64)
    shift = (off[g_arm] - local_off[s_arm]) if len(s_arm) else np.zeros(0, np.int64)
    cols = [g_arm.astype(np.int64), (s_lo + shift).astype(np.int64), (s_hi + shift).astype(np.int64),
            s_mean.astype(np.float64)]
    full = gather_tables(cols, dst=dst, group=group, device=device)
    if full is None:
        return None
    order = np.lexsort((full[1], full[0]))
    return tuple(c[order] for c in full)
