"""Device-resident truth index and the top-n search over it (thin object layer over the C ABI).

`TruthIndex` owns one `ds_index` (one shard of the truth DB on one GPU).  Inputs may be numpy arrays
(host buffers: staged by the library, outputs come back as numpy) or CUDA torch tensors (used in
place, outputs are CUDA tensors and the call is asynchronous on torch's current stream).
"""
import ctypes

import numpy as np

from . import _native as nat


def _is_cuda_tensor(x):
    return hasattr(x, 'is_cuda') and x.is_cuda


def _empty(shape, dtype, like_cuda, device=None):
    if like_cuda:
        import torch
        return torch.empty(shape, dtype=getattr(torch, dtype), device=device)
    return np.empty(shape, dtype=dtype)


class TruthIndex:
    """Truth rows of one shard packed on one GPU (replaces the truth half of MatchMaker.__init__,
    /root/reference/doppelspeller/match_maker.py:97-109)."""

    def __init__(self, t_row_ptr, t_col_ids, idf64_by_col, sums=None, device=0, row_offset=0, n_total=None):
        nat.expect(t_row_ptr, 'int64', 't_row_ptr')
        nat.expect(t_col_ids, 'uint16', 't_col_ids')
        nat.expect(idf64_by_col, 'float64', 'idf64_by_col')
        if sums is not None:
            nat.expect(sums, 'float32', 'sums')
        self.device = int(device)
        self.n_truth = int(t_row_ptr.shape[0]) - 1
        self.n_vocab = int(idf64_by_col.shape[0])
        self.row_offset = int(row_offset)
        self.n_total = int(n_total) if n_total is not None else self.n_truth
        handle = ctypes.c_void_p()
        nat.check(nat.lib.ds_index_create(
            ctypes.byref(handle), self.device, self.n_truth, self.n_vocab, nat.ptr(t_row_ptr), nat.ptr(t_col_ids),
            nat.ptr(idf64_by_col), nat.ptr(sums), self.row_offset, self.n_total, self._stream()))
        self._handle = handle

    def _stream(self):
        import torch
        if not torch.cuda.is_available():
            return None
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, '_handle', None):
            nat.lib.ds_index_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sums(self):
        out = np.empty(self.n_truth, dtype=np.float32)
        nat.check(nat.lib.ds_index_get_sums(self._handle, nat.ptr(out), self._stream()))
        return out

    @staticmethod
    def _check_queries(q_row_ptr, q_col_ids, q_mx):
        nat.expect(q_row_ptr, 'int64', 'q_row_ptr')
        nat.expect(q_col_ids, 'uint16', 'q_col_ids')
        if q_mx is not None:
            nat.expect(q_mx, 'float64', 'q_mx')

    def topn(self, q_row_ptr, q_col_ids, k, q_mx=None, mx_mode=nat.DS_MX_PY312_COMPENSATED, with_details=False,
             out_rows=None, out_count=None):
        """`[get_closest_matches(q) for q in queries]` as truth ROW indexes (match_maker.py:192-203).

        Returns (rows int64[Q,k] descending row index, count int32[Q]) and, with_details, also
        (kth_key float32[Q], flags int32[Q]).  `out_rows` / `out_count` may be preallocated buffers (e.g.
        pinned host tensors) to receive the result."""
        self._check_queries(q_row_ptr, q_col_ids, q_mx)
        n_q = int(q_row_ptr.shape[0]) - 1
        cuda = _is_cuda_tensor(q_col_ids)
        dev = q_col_ids.device if cuda else None
        rows = out_rows if out_rows is not None else _empty((n_q, k), 'int64', cuda, dev)
        count = out_count if out_count is not None else _empty((n_q,), 'int32', cuda, dev)
        nat.expect(rows, 'int64', 'out_rows')
        nat.expect(count, 'int32', 'out_count')
        if tuple(rows.shape) != (n_q, k) or tuple(count.shape) != (n_q,):
            raise ValueError('out_rows / out_count have the wrong shape')
        kth = _empty((n_q,), 'float32', cuda, dev) if with_details else None
        flags = _empty((n_q,), 'int32', cuda, dev) if with_details else None
        nat.check(nat.lib.ds_topn(self._handle, n_q, nat.ptr(q_row_ptr), nat.ptr(q_col_ids), nat.ptr(q_mx), mx_mode, k,
                                  nat.ptr(rows), nat.ptr(count), nat.ptr(kth), nat.ptr(flags), self._stream()))
        if with_details:
            return rows, count, kth, flags
        return rows, count

    # ---- sharded phases ----
    def topn_local(self, q_row_ptr, q_col_ids, k, q_mx=None, mx_mode=nat.DS_MX_PY312_COMPENSATED, theta_own=None, theta_peers=()):
        """Phase 1: this shard's best m = topn_retained(k) candidates per query -> (score f64[Q,m],
        global row int64[Q,m], mx f64[Q]).  `theta_own` (float64 CUDA tensor [Q], zeroed and synchronised by the
        caller) and `theta_peers` (peer-mapped device addresses of the other shards' arrays) share the pruning
        thresholds between the shards while they scan (ds_topn_local_shared)."""
        self._check_queries(q_row_ptr, q_col_ids, q_mx)
        n_q = int(q_row_ptr.shape[0]) - 1
        m = nat.topn_retained(k)
        cuda = _is_cuda_tensor(q_col_ids)
        dev = q_col_ids.device if cuda else None
        score = _empty((n_q, m), 'float64', cuda, dev)
        row = _empty((n_q, m), 'int64', cuda, dev)
        mx = _empty((n_q,), 'float64', cuda, dev)
        if theta_own is None:
            nat.check(nat.lib.ds_topn_local(self._handle, n_q, nat.ptr(q_row_ptr), nat.ptr(q_col_ids), nat.ptr(q_mx), mx_mode, k,
                                            nat.ptr(score), nat.ptr(row), nat.ptr(mx), self._stream()))
        else:
            nat.expect(theta_own, 'float64', 'theta_own')
            peers = (ctypes.c_void_p * max(1, len(theta_peers)))(*[int(p) for p in theta_peers])
            nat.check(nat.lib.ds_topn_local_shared(self._handle, n_q, nat.ptr(q_row_ptr), nat.ptr(q_col_ids), nat.ptr(q_mx), mx_mode, k,
                                                   nat.ptr(score), nat.ptr(row), nat.ptr(mx), nat.ptr(theta_own), peers,
                                                   len(theta_peers), self._stream()))
        return score, row, mx

    def topn_rescan(self, q_row_ptr, q_col_ids, q_mx, threshold, flags, k, rows, count):
        """Phase 3: patches `rows` / `count` of the flagged queries with the k highest local rows whose
        exact score reaches `threshold`."""
        self._check_queries(q_row_ptr, q_col_ids, q_mx)
        n_q = int(q_row_ptr.shape[0]) - 1
        nat.check(nat.lib.ds_topn_rescan(self._handle, n_q, nat.ptr(q_row_ptr), nat.ptr(q_col_ids), nat.ptr(q_mx),
                                         nat.ptr(threshold), nat.ptr(flags), k, nat.ptr(rows), nat.ptr(count), self._stream()))
        return rows, count


def topn_merge(all_score, all_row, k, n_total, q_mx=None, device=0):
    """Phase 2 over the gathered [n_shards, Q, m] candidates -> (rows, count, kth_key, threshold, flags)."""
    n_shards, n_q, m = all_score.shape
    if m != nat.topn_retained(k):
        raise ValueError(f'candidate lists hold {m} slots, expected {nat.topn_retained(k)} for k={k}')
    cuda = _is_cuda_tensor(all_score)
    dev = all_score.device if cuda else None
    rows = _empty((n_q, k), 'int64', cuda, dev)
    count = _empty((n_q,), 'int32', cuda, dev)
    kth = _empty((n_q,), 'float32', cuda, dev)
    thr = _empty((n_q,), 'float64', cuda, dev)
    flags = _empty((n_q,), 'int32', cuda, dev)
    nat.check(nat.lib.ds_topn_merge(n_shards, n_q, k, n_total, nat.ptr(all_score), nat.ptr(all_row), nat.ptr(q_mx),
                                    nat.ptr(rows), nat.ptr(count), nat.ptr(kth), nat.ptr(thr), nat.ptr(flags),
                                    int(dev.index) if cuda else int(device), nat.stream_for(all_score)))
    return rows, count, kth, thr, flags
