"""CPU, world_size 2 over gloo: the multi-rank plumbing of doppelspeller_b200/sharded.py (all_gather of
the phase-1 lists, replicated merge, rescan exchange, highest-shard-first combination) with an
oracle-backed shard standing in for the GPU kernels.  The kernels themselves are covered on the GPU by
tests/test_gpu_parity.py::test_sharded_phases_on_one_gpu_match_single_index."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class OracleShard:
    """Implements local / merge / rescan of one shard with the CPU oracle's exact scores (float64),
    following the phase contracts of include/doppelspeller_b200.h."""

    def __init__(self, index, r0, r1, n_total, retained):
        from oracle import oracle
        self.oracle, self.index, self.r0, self.r1, self.n_total, self.m = oracle, index, r0, r1, n_total, retained
        self.n_q = int(index['qs_ptr'].shape[0]) - 1
        self.scores = np.stack([oracle.fast_jaccard(index, q)[r0:r1] for q in range(self.n_q)])

    def local(self, k):
        score = np.full((self.n_q, self.m), -1.0)
        row = np.full((self.n_q, self.m), -1, dtype=np.int64)
        for q in range(self.n_q):
            s = self.scores[q]
            pos = np.nonzero(s > 0)[0]
            order = pos[np.lexsort((-pos, -s[pos]))][:self.m]          # score desc, ties: higher row first
            score[q, :len(order)] = s[order]
            row[q, :len(order)] = order + self.r0
        mx = np.array([self.oracle.py_float_sum(self.index['w64'], self.index['qs_cols'][self.index['qs_ptr'][q]:self.index['qs_ptr'][q + 1]])
                       for q in range(self.n_q)])
        return torch.as_tensor(score), torch.as_tensor(row), torch.as_tensor(mx)

    def merge(self, all_score, all_row, k, q_mx):
        all_score, all_row = all_score.numpy(), all_row.numpy()
        n_shards = all_score.shape[0]
        rows = np.full((self.n_q, k), -1, dtype=np.int64)
        count = np.zeros(self.n_q, dtype=np.int32)
        kth = np.zeros(self.n_q, dtype=np.float32)
        thr = np.zeros(self.n_q)
        flags = np.zeros(self.n_q, dtype=np.int32)
        for q in range(self.n_q):
            valid = all_row[:, q, :] >= 0
            s, r = all_score[:, q, :][valid], all_row[:, q, :][valid]
            if len(s) >= k:
                kth[q] = np.float32(np.sort(s)[::-1][k - 1])
            thr[q] = float(kth[q]) - float(np.float32(1e-6))
            if not thr[q] > 0:
                flags[q] = 2
                c = min(k, self.n_total)
                rows[q, :c] = self.n_total - 1 - np.arange(c)
                count[q] = c
                continue
            incomplete = any(all_row[sh, q, self.m - 1] >= 0 and all_score[sh, q, self.m - 1] >= thr[q] for sh in range(n_shards))
            if incomplete:
                flags[q] = 1
                continue
            qual = np.sort(r[s >= thr[q]])[::-1][:k]
            rows[q, :len(qual)] = qual
            count[q] = len(qual)
        return (torch.as_tensor(rows), torch.as_tensor(count), torch.as_tensor(kth), torch.as_tensor(thr), torch.as_tensor(flags))

    def rescan(self, q_mx, threshold, flags, k):
        rows = np.full((self.n_q, k), -1, dtype=np.int64)
        count = np.zeros(self.n_q, dtype=np.int32)
        for q in np.nonzero(flags.numpy() & 1)[0]:
            qual = np.nonzero(self.scores[q] >= float(threshold[q]))[0][::-1][:k]
            rows[q, :len(qual)] = qual + self.r0
            count[q] = len(qual)
        return torch.as_tensor(rows), torch.as_tensor(count)


def _worker(rank, world, port, k, retained, out_path):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from doppelspeller_b200 import sharded
    from tests.test_sharded_gloo import OracleShard, make_case
    index, n_total = make_case()
    offs = sharded.shard_offsets(n_total, world)
    shard = OracleShard(index, int(offs[rank]), int(offs[rank + 1]), n_total, retained)
    rows, count, flags = sharded.sharded_topn(shard, k)
    if rank == 0:
        np.savez(out_path, rows=rows.numpy(), count=count.numpy(), flags=flags.numpy())
    dist.barrier()
    dist.destroy_process_group()


def make_case():
    from oracle import oracle
    rng = np.random.default_rng(3)
    n_total, n_vocab, n_q = 600, 50, 40
    truth = [sorted(rng.choice(n_vocab, size=int(rng.integers(2, 7)), replace=False).tolist()) for _ in range(n_total - 120)]
    truth += [[0, 1, 2]] * 120                                     # massive ties -> retained lists overflow -> rescan
    order = rng.permutation(n_total)
    truth = [truth[i] for i in order]
    queries = [sorted(rng.choice(n_vocab, size=int(rng.integers(1, 8)), replace=False).tolist()) for _ in range(n_q - 3)]
    queries += [[0, 1, 2], [0, 1], []]
    df = np.bincount(np.concatenate([np.array(t) for t in truth]), minlength=n_vocab)
    w64 = np.array([np.log(n_total / d) if d > 0 else 0.0 for d in df])
    w64[df == 0] = w64.max()

    def csr(sets):
        ptr = np.zeros(len(sets) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in sets], out=ptr[1:])
        return ptr, np.array([c for s in sets for c in s], dtype=np.int32)
    t_ptr, t_cols = csr(truth)
    q_ptr, q_cols = csr(queries)
    index = oracle.finish_index(dict(n_truth=n_total, w64=w64, w32=w64.astype(np.float32), t_ptr=t_ptr, t_cols=t_cols,
                                     q_ptr=q_ptr, q_cols=q_cols))
    return index, n_total


@pytest.mark.parametrize('k,retained', [(5, 16), (10, 32)])
def test_two_rank_sharded_topn_matches_oracle(tmp_path, k, retained):
    from oracle import oracle
    out_path = str(tmp_path / 'result.npz')
    mp.spawn(_worker, args=(2, _free_port(), k, retained, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    index, n_total = make_case()
    want_rows, want_count, _ = oracle.topn(index, k)
    assert np.array_equal(got['count'], want_count)
    assert np.array_equal(got['rows'], want_rows)
    assert (got['flags'] & 1).any(), 'the case is meant to exercise the rescan exchange'


def test_shard_offsets_and_slicing():
    from doppelspeller_b200 import sharded
    offs = sharded.shard_offsets(10, 3)
    assert offs.tolist() == [0, 4, 7, 10]
    ptr = np.array([0, 2, 2, 5, 6], dtype=np.int64)
    cols = np.arange(6, dtype=np.uint16)
    p, c = sharded.slice_truth_csr(ptr, cols, 1, 3)
    assert p.tolist() == [0, 0, 3] and c.tolist() == [2, 3, 4]
