"""Top-n search over a truth DB sharded across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r holds the contiguous truth rows
[offsets[r], offsets[r+1]) as its own `TruthIndex`; the IDF weights are global (computed over the
whole DB before sharding) and the queries are replicated.  The reference's selection rule needs one
global quantity in the middle - the k-th largest float32 score - so the path has exactly one
exchange step:

    phase 1  local scan           -> best m candidates per query  (score64, global row)
    all_gather of [Q, m] scores + rows over the shards            (the only collective on the path)
    phase 2  merge (replicated)   -> global k-th key, threshold, final rows or a RESCAN flag
    phase 3  rescan (flagged only) + all_gather of the [F, k] local answers, highest shard first

Without sharding (one rank) no collective is issued at all.
"""
import numpy as np

from . import _native as nat


def shard_offsets(n_total, n_shards):
    """Contiguous ascending row ranges: shard r owns [off[r], off[r+1])."""
    base, extra = divmod(int(n_total), int(n_shards))
    sizes = [base + (1 if r < extra else 0) for r in range(n_shards)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def slice_truth_csr(t_row_ptr, t_col_ids, r0, r1):
    """Rows [r0, r1) of a CSR as their own CSR (numpy arrays or torch tensors, host or device)."""
    ptr = t_row_ptr[r0:r1 + 1]
    cols = t_col_ids[int(ptr[0]):int(ptr[-1])]
    if isinstance(ptr, np.ndarray):
        return np.ascontiguousarray(ptr - ptr[0]), np.ascontiguousarray(cols)
    return (ptr - ptr[0]).contiguous(), cols.contiguous()


def combine_rescans(per_shard_rows, per_shard_count, k):
    """Final rows of the re-scanned queries: every shard reports its k highest qualifying rows
    (descending); shards are contiguous ascending ranges, so the answer is the concatenation from the
    highest shard down, cut at k.  per_shard_rows [S, F, k], per_shard_count [S, F] (torch tensors).
    Fully vectorised (no host round trips)."""
    import torch
    n_shards, n_f, _ = per_shard_rows.shape
    rows = per_shard_rows.flip(0).permute(1, 0, 2).reshape(n_f, n_shards * k)          # highest shard first
    counts = per_shard_count.flip(0).to(torch.int64).t()                               # [F, S]
    slot = torch.arange(k, device=rows.device).view(1, 1, k)
    valid = (slot < counts.unsqueeze(2)).reshape(n_f, n_shards * k)
    position = torch.cumsum(valid.to(torch.int64), dim=1) - 1
    keep = valid & (position < k)
    out = torch.full((n_f, k), -1, dtype=torch.int64, device=rows.device)
    f_index = torch.arange(n_f, device=rows.device).unsqueeze(1).expand_as(rows)
    out[f_index[keep], position[keep]] = rows[keep]
    filled = torch.clamp(counts.sum(dim=1), max=k).to(torch.int32)
    return out, filled


_SHARED_THETA = {}


def shared_thresholds(n_q, device, group):
    """Symmetric (NVLink peer-mapped) float64[n_q] buffer of this rank + the peers' addresses, for ds_topn_local_shared.
    Allocated and exchanged once per (size, device, group) - the rendezvous is a collective.  None when symmetric
    memory is not available (single rank, CPU backends, no peer access)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    key = (int(n_q), int(device.index), id(group))
    if key not in _SHARED_THETA:
        entry = None
        try:
            import torch.distributed._symmetric_memory as symm
            theta = symm.empty(max(1, int(n_q)), dtype=torch.float64, device=device)
            handle = symm.rendezvous(theta, group if group is not None else dist.group.WORLD)
            peers = [int(handle.buffer_ptrs[r]) for r in range(handle.world_size) if r != handle.rank]
            if len(peers) <= 31:
                entry = (theta, handle, peers)
        except Exception as error:   # the scan then runs with local thresholds only
            import warnings
            warnings.warn(f'shared thresholds unavailable: {error!r}')
        _SHARED_THETA[key] = entry
    return _SHARED_THETA[key]


class GpuShard:
    """Kernel backend of one rank: wraps a TruthIndex, moves the replicated queries to its device once.
    `group` (the ranks holding the other shards of the same truth DB) enables the shared pruning thresholds."""

    def __init__(self, index, q_row_ptr, q_col_ids, mx_mode=nat.DS_MX_PY312_COMPENSATED, group=None, share_thresholds=False):
        import torch
        self.index = index
        self.device = torch.device('cuda', index.device)
        self._shared = shared_thresholds(int(q_row_ptr.shape[0]) - 1, self.device, group) if share_thresholds else None
        if isinstance(q_row_ptr, np.ndarray):
            q_row_ptr = torch.as_tensor(np.ascontiguousarray(q_row_ptr, dtype=np.int64))
            q_col_ids = torch.as_tensor(np.ascontiguousarray(q_col_ids, dtype=np.uint16))
        # pinned host tensors are copied asynchronously on the current stream (e2e path)
        self.q_ptr = q_row_ptr.to(self.device, non_blocking=True).contiguous()
        self.q_cols = q_col_ids.to(self.device, non_blocking=True).contiguous()
        self.mx_mode = mx_mode
        self.n_total = index.n_total

    def local(self, k):
        if self._shared is None:
            return self.index.topn_local(self.q_ptr, self.q_cols, k, mx_mode=self.mx_mode)
        theta, handle, peers = self._shared
        theta.zero_()
        handle.barrier()      # every shard's zeros are in place before any shard reads them (device side, this stream)
        return self.index.topn_local(self.q_ptr, self.q_cols, k, mx_mode=self.mx_mode, theta_own=theta, theta_peers=peers)

    def merge(self, all_score, all_row, k, q_mx):
        from .index import topn_merge
        return topn_merge(all_score, all_row, k, self.n_total, q_mx=q_mx, device=self.index.device)

    def rescan(self, q_mx, threshold, flags, k):
        import torch
        n_q = self.q_ptr.shape[0] - 1
        rows = torch.full((n_q, k), -1, dtype=torch.int64, device=self.device)
        count = torch.zeros(n_q, dtype=torch.int32, device=self.device)
        self.index.topn_rescan(self.q_ptr, self.q_cols, q_mx, threshold, flags, k, rows, count)
        return rows, count


def _all_gather(tensor, group, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return tensor.unsqueeze(0)
    tensor = tensor.contiguous()
    if tensor.dim() == 0 or tensor.shape[0] == 0:
        parts = [torch.empty_like(tensor) for _ in range(world)]
        dist.all_gather(parts, tensor, group=group)
        return torch.stack(parts)
    flat = torch.empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(flat, tensor, group=group)       # concatenated along dim 0 (NCCL and gloo)
    return flat.view((world,) + tuple(tensor.shape))


def sharded_topn(shard, k, group=None, timings=None):
    """Runs the three phases on this rank's `shard` (GpuShard or any object with local / merge / rescan).
    Returns (rows int64[Q,k] global truth rows in descending order, count int32[Q], flags int32[Q]) -
    identical on every rank.  `timings` (dict) receives per-phase milliseconds when given (CUDA tensors only)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    marks = []

    def mark(name):
        if timings is not None and torch.cuda.is_available():
            event = torch.cuda.Event(enable_timing=True)
            event.record()
            marks.append((name, event))

    mark('start')
    score, row, mx = shard.local(k)
    mark('local')
    all_score = _all_gather(score, group, world)
    all_row = _all_gather(row, group, world)
    mark('all_gather')
    rows, count, _, threshold, flags = shard.merge(all_score, all_row, k, mx)
    mark('merge')
    flagged = torch.nonzero((flags & nat.DS_FLAG_RESCAN) != 0).flatten()
    if flagged.numel() > 0:          # the same set on every rank: merge inputs are replicated
        local_rows, local_count = shard.rescan(mx, threshold, flags, k)
        mark('rescan')
        per_rows = _all_gather(local_rows[flagged], group, world)
        per_count = _all_gather(local_count[flagged], group, world)
        fixed_rows, fixed_count = combine_rescans(per_rows, per_count, k)
        rows[flagged] = fixed_rows
        count[flagged] = fixed_count
        mark('rescan_exchange')
    if timings is not None and marks:
        marks[-1][1].synchronize()
        for (_, before), (name, after) in zip(marks[:-1], marks[1:]):
            timings[name] = timings.get(name, 0.0) + before.elapsed_time(after)
    return rows, count, flags
