"""CPU models of two invariants the posting-list kernel (csrc/ds_topn.cu k_post) rests on - checked in numpy so that they
are guarded where no GPU is present:

  1. the float32 pre-filter `sc > fmaf(a, sums, b)` derived from a float64 threshold never rejects a pair whose
     reference score reaches the threshold (filter_from_threshold), and a group's bar `fmaf(a, min sums of the group, b)`
     never exceeds the bar of any of its rows;
  2. the bank-balanced order of a segment (k_post_balance) is a permutation that puts 32 consecutive postings into
     distinct shared-memory banks whenever the segment allows it, and walking blocks / pieces in that order adds the
     same float32 values in the same per-row order as the reference's column-by-column accumulation.
"""
import numpy as np


def _f32_round_down(x64):
    """__double2float_rd"""
    y = np.asarray(x64, dtype=np.float64).astype(np.float32)
    too_big = y.astype(np.float64) > x64
    return np.where(too_big, np.nextafter(y, np.float32(-np.inf)), y).astype(np.float32)


def _fmaf(a, x, b):
    """float32 fused multiply-add: the product of two float32 is exact in float64; one rounding to float32 at the end
    (the float64 addition can round first, which is immaterial at the margins tested here)."""
    return (np.asarray(a, np.float64) * np.asarray(x, np.float64) + np.asarray(b, np.float64)).astype(np.float32)


def _filter_from_threshold(theta, mx):
    """ds_topn.cu filter_from_threshold"""
    c = theta / (1.0 + theta)
    a = _f32_round_down(c * (1.0 - 4e-6))
    b = _f32_round_down(a.astype(np.float64) * mx * (1.0 - 1e-6))
    return a, np.maximum(b, np.float32(0.0))


def test_prefilter_has_no_false_negatives():
    rng = np.random.default_rng(7)
    n = 400_000
    mx = rng.uniform(5.0, 400.0, n)                                        # float64 sum of the query's idf weights
    sums = rng.uniform(1.0, 400.0, n).astype(np.float32)                   # float32 sum of the truth row's weights
    sc = (np.minimum(sums.astype(np.float64), mx) * rng.uniform(0.0, 1.0, n)).astype(np.float32)
    s64 = sc.astype(np.float64) / (sums.astype(np.float64) + (mx - sc.astype(np.float64)))   # match_maker.py:50
    # thresholds at, just below and just above the pair's own score: the hardest cases for a conservative filter
    for theta in (s64, s64 * (1 - 1e-12), np.nextafter(s64, 0.0), rng.uniform(0.01, 0.9, n)):
        a, b = _filter_from_threshold(theta, mx)
        passes = sc > _fmaf(a, sums, b)
        qualifies = (s64 >= theta) & (s64 > 0)
        assert not (qualifies & ~passes).any()
        # a group's bar uses the smallest row sum of the group: never above the row's own bar
        floor = np.minimum(sums, rng.uniform(0.5, 400.0, n).astype(np.float32))
        assert (_fmaf(a, floor, b) <= _fmaf(a, sums, b)).all()
        assert (_fmaf(a, floor, b) >= b).all()


def _balanced_positions(rows):
    """k_post_balance: rank by (index within the bank, bank); the index within a bank is any bijection (atomics)."""
    bank = rows & 31
    counts = np.bincount(bank, minlength=32)
    k = np.zeros(len(rows), dtype=np.int64)
    seen = np.zeros(32, dtype=np.int64)
    for i, b in enumerate(bank):
        k[i] = seen[b]
        seen[b] += 1
    at = np.minimum(counts[None, :], k[:, None]).sum(1) + ((np.arange(32)[None, :] < bank[:, None]) & (counts[None, :] > k[:, None])).sum(1)
    return at


def test_bank_balanced_order():
    rng = np.random.default_rng(11)
    for n in (33, 64, 200, 777, 2048):
        rows = rng.choice(2048, n, replace=False)
        at = _balanced_positions(rows)
        assert sorted(at.tolist()) == list(range(n))                         # a permutation
        ordered = np.empty(n, dtype=np.int64)
        ordered[at] = rows
        fullest = np.bincount(rows & 31, minlength=32).min()                # every bank has at least this many postings
        for start in range(0, fullest * 32, 32):                            # ... so these slabs are conflict free
            assert len(set((ordered[start:start + 32] & 31).tolist())) == 32


def test_block_and_piece_walk_reproduces_the_reference_accumulation():
    from doppelspeller_b200 import encode, synthetic
    truth = synthetic.generate_truth_titles(9000, seed=5)
    test, _ = synthetic.generate_test_titles(truth, 12, seed=6)
    enc = encode.encode_canonical(test, truth)
    w32 = enc['idf64'].astype(np.float32)
    t_ptr, t_cols = enc['t_ptr'], enc['t_cols'].astype(np.int64)
    n, block = len(truth), 2048
    row_of = np.repeat(np.arange(n), np.diff(t_ptr))
    for q in range(12):
        cols = np.sort(enc['q_cols'][enc['q_ptr'][q]:enc['q_ptr'][q + 1]].astype(np.int64))
        reference = np.zeros(n, dtype=np.float32)                           # match_maker.py:33-47
        for c in cols:
            rows = row_of[t_cols == c]
            reference[rows] = reference[rows] + w32[c]
        walked = np.zeros(n, dtype=np.float32)
        for first in range(0, n, block):                                     # one warp task = one block at a time
            acc = np.zeros(block, dtype=np.float32)
            for c in cols:                                                   # ascending column ids
                rows = row_of[(t_cols == c) & (row_of >= first) & (row_of < first + block)] - first
                if len(rows) > 32:
                    order = np.empty(len(rows), dtype=np.int64)
                    order[_balanced_positions(rows)] = rows
                    rows = order
                for piece in range(0, len(rows), 64):                        # <= 64 postings per piece
                    part = rows[piece:piece + 64]
                    acc[part] = acc[part] + w32[c]
            walked[first:first + block] = acc[:min(block, n - first)]
        assert np.array_equal(walked.view(np.uint32), reference.view(np.uint32))
