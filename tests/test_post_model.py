"""CPU models of the invariants the posting-list kernel (csrc/ds_topn.cu k_post) rests on - checked in numpy so that
they are guarded where no GPU is present:

  1. the float32 pre-filter `sc > fmaf(a, sums, b)` derived from a float64 threshold never rejects a pair whose
     reference score reaches the threshold (filter_from_threshold), and a group's bar `fmaf(a, min sums of the group, b)`
     never exceeds the bar of any of its rows;
  2. the bank-balanced order of a segment (k_post_balance) is a permutation that puts 32 consecutive postings into
     distinct shared-memory banks (u16 accumulators: bank = (row / 2) & 31) whenever the segment allows it;
  3. k_post's conservative test (`post_test`: 16-bit fixed-point sum of the sparse columns, weights rounded up, plus the
     float32 gain of the dense columns, times `grow`) passes whenever the reference's own float32 score sum passes the
     pre-filter, no accumulator overflows, and a group's bar (largest accumulator value failing the test with the
     group's smallest row sum and OR-ed dense pattern) is never reached by a row that passes its own test.
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
    bank = (rows >> 1) & 31
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
    for n in (33, 64, 200, 777, 2048, 4096):
        rows = rng.choice(4096, n, replace=False)
        at = _balanced_positions(rows)
        assert sorted(at.tolist()) == list(range(n))                         # a permutation
        ordered = np.empty(n, dtype=np.int64)
        ordered[at] = rows
        fullest = np.bincount((rows >> 1) & 31, minlength=32).min()         # every bank has at least this many postings
        for start in range(0, fullest * 32, 32):                            # ... so these slabs are conflict free
            assert len(set(((ordered[start:start + 32] >> 1) & 31).tolist())) == 32


# ---- directed float32 rounding of a float64 value (the products / sums below are exact or nearly so in float64) ----
def _ru(x64):
    y = np.asarray(x64, dtype=np.float64).astype(np.float32)
    return np.where(y.astype(np.float64) < x64, np.nextafter(y, np.float32(np.inf)), y).astype(np.float32)


def _post_test(acc, gain, bar, inv_scale, grow):
    """ds_topn.cu post_test: fmul_ru(fadd_ru(fmul_ru(acc, inv_scale), gain), grow) > bar"""
    t = _ru(np.float64(acc) * np.float64(inv_scale))
    t = _ru(np.float64(t) + np.float64(gain))
    t = _ru(np.float64(t) * np.float64(grow))
    return t > bar


def _dense_gain(bits, dense_w):
    v = np.float32(0.0)
    for bit in range(32):
        if (bits >> bit) & 1:
            v = np.float32(v + dense_w[bit])
    return v


def test_fixed_point_test_has_no_false_negatives():
    rng = np.random.default_rng(23)
    misses = 0
    for trial in range(1500):
        g = int(rng.integers(1, 60 if trial % 10 else 400))
        n_dense = int(rng.integers(0, min(g, 20) + 1))
        w = np.sort(rng.uniform(0.05, 13.0, g)).astype(np.float32)[::-1].copy()       # any order: column ids are random below
        if trial % 7 == 0:
            w[:] = np.float32(rng.uniform(0.5, 12.0))                                 # equal weights: rounding worst cases
        col = rng.permutation(g)                                                      # ascending column id = reference order
        is_dense = np.zeros(g, dtype=bool)
        is_dense[rng.choice(g, n_dense, replace=False)] = True
        bit_of = np.full(g, -1)
        bit_of[is_dense] = rng.choice(32, n_dense, replace=False)
        dense_w = np.zeros(32, dtype=np.float32)
        dense_w[bit_of[is_dense]] = w[is_dense]
        mx = np.float64(w.astype(np.float64).sum())
        # k_post's per query constants
        sparse_sum = np.float32(0.0)
        for x in w[~is_dense]:
            sparse_sum = _ru(np.float64(sparse_sum) + np.float64(x))
        if sparse_sum > 0:
            scale = np.float32(np.float64(65535 - 4 - 2 * min(g, 16000)) / np.float64(sparse_sum))
            if np.float64(scale) * np.float64(sparse_sum) > 65535 - 4 - 2 * min(g, 16000):
                scale = np.nextafter(scale, np.float32(0))                            # __fdiv_rd
            inv_scale = _ru(1.0 / np.float64(scale))
        else:
            scale, inv_scale = np.float32(0.0), np.float32(0.0)
        grow = _ru(1.0 + np.float64(_ru(np.float64(2 * g + 64) * np.float64(np.float32(5.9604645e-8)))))
        fix = np.minimum(65535, np.ceil(np.atleast_1d(_ru(w.astype(np.float64) * np.float64(scale))).astype(np.float64))).astype(np.int64)
        assert fix[~is_dense].sum() <= 65535
        qmask = int(sum(1 << int(b) for b in bit_of[is_dense]))
        for _ in range(40):
            hit = rng.random(g) < rng.uniform(0.05, 1.0)
            order = np.argsort(col)
            sc = np.float32(0.0)
            for i in order:                                                           # match_maker.py:45-48
                if hit[i]:
                    sc = np.float32(sc + w[i])
            sums = np.float32(max(float(sc), rng.uniform(1.0, 300.0)))
            s64 = np.float64(sc) / (np.float64(sums) + (mx - np.float64(sc)))
            for theta in (s64, np.nextafter(s64, 0.0), s64 * 0.999, rng.uniform(0.01, 0.9)):
                if not theta > 0:
                    continue
                a, b = _filter_from_threshold(np.float64(theta), mx)
                bar = _fmaf(a, sums, b)
                reference_passes = sc > bar
                acc = int(fix[hit & ~is_dense].sum())
                pattern = int(sum(1 << int(bit_of[i]) for i in range(g) if hit[i] and is_dense[i]))
                gain = _dense_gain(pattern & qmask, dense_w)
                ours = bool(_post_test(acc, gain, bar, inv_scale, grow))
                if reference_passes and not ours:
                    misses += 1
                # group bar: smaller row sum, larger pattern -> the largest failing accumulator value
                floor = np.float32(sums * rng.uniform(0.6, 1.0))
                union = pattern | int(rng.integers(0, 1 << 32)) & qmask
                g_gain = _dense_gain(union & qmask, dense_w)
                g_bar = _fmaf(a, floor, b)
                room = (np.float64(g_bar) * (2.0 - np.float64(grow)) - np.float64(g_gain)) * np.float64(scale)
                guess = 65534 if room >= 65534 else (int(room) - 1 if room >= 1 else -1)
                if inv_scale == 0:
                    guess = -1 if _post_test(0, g_gain, g_bar, inv_scale, grow) else 65535
                if guess < 0 and not _post_test(0, g_gain, g_bar, inv_scale, grow):
                    guess = 0
                if 0 <= guess < 65535:
                    it = 0
                    while it < 64 and guess < 65535 and not _post_test(guess + 1, g_gain, g_bar, inv_scale, grow):
                        guess += 1
                        it += 1
                    it = 0
                    while it < 64 and guess >= 0 and _post_test(guess, g_gain, g_bar, inv_scale, grow):
                        guess -= 1
                        it += 1
                    if guess >= 0 and _post_test(guess, g_gain, g_bar, inv_scale, grow):
                        guess = -1
                if ours:
                    assert acc > guess, (trial, acc, guess)
    assert misses == 0
