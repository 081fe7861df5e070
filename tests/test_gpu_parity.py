"""GPU: parity of the CUDA path (through the C ABI) with the golden vectors minted from the reference
and with the CPU oracle on seeded inputs.  Bit-exact for rows / distances / integer features; 1e-6
relative for the float32 idf / rank features (north_star tolerance)."""
import numpy as np
import pytest

from tests.conftest import features_equal, oracle_index_from_encoded

pytestmark = pytest.mark.gpu


def _match_maker(enc, k, title_ids=None, **kw):
    from doppelspeller_b200.match_maker import MatchMaker
    n_truth = int(enc['t_ptr'].shape[0]) - 1
    ids = np.arange(n_truth) if title_ids is None else title_ids
    return MatchMaker.from_encoded(enc['idf64'], enc['t_ptr'], enc['t_cols'], enc['q_ptr'], enc['q_cols'], ids, k, **kw)


def _golden_enc(g):
    return dict(idf64=g['w64'], t_ptr=g['t_ptr'], t_cols=g['t_cols'], q_ptr=g['q_ptr'], q_cols=g['q_cols'])


# ------------------------------------------------------------------ K1: candidates
def test_sums_match_reference(golden_matchmaker):
    mm = _match_maker(_golden_enc(golden_matchmaker), 10)
    assert np.array_equal(mm.sums_matrix_truth.view(np.uint32), golden_matchmaker['sums'].view(np.uint32))


@pytest.mark.parametrize('k,key', [(10, 'top10_rows'), (100, 'top100_rows')])
def test_candidates_match_reference_example(golden_matchmaker, k, key):
    g = golden_matchmaker
    mm = _match_maker(_golden_enc(g), k, title_ids=g['title_ids'])
    rows, count = mm.closest_rows()
    assert (count == k).all()
    assert np.array_equal(rows, g[key].astype(np.int64))
    # the drop-in call: title ids in the reference's order
    for q in (0, 1, 17, 999):
        assert mm.get_closest_matches(q) == g['title_ids'][g[key][q]].tolist()


@pytest.mark.parametrize('n_truth,n_q,k,seed', [(20000, 1500, 10, 1), (20000, 700, 100, 2), (3000, 300, 1, 3),
                                                 (50000, 400, 10, 4), (700, 100, 512, 5)])
def test_candidates_match_oracle_synthetic(n_truth, n_q, k, seed):
    from doppelspeller_b200 import encode, synthetic
    from oracle import oracle
    truth = synthetic.generate_truth_titles(n_truth, seed=seed)
    test, _ = synthetic.generate_test_titles(truth, n_q, seed=seed + 100)
    enc = encode.encode_canonical(test, truth)
    rows, count, kth, flags = _match_maker(enc, k)._index.topn(enc['q_ptr'], enc['q_cols'], k, with_details=True)
    want_rows, want_count, want_kth = oracle.topn(oracle_index_from_encoded(enc), k)
    assert np.array_equal(count, want_count)
    assert np.array_equal(rows, want_rows)
    assert np.array_equal(kth, want_kth)


def test_negative_weight_keeps_the_dense_scan():
    """The posting-list kernel needs partial sums that only grow; an index with a negative idf weight must stay on
    the dense row scan and still reproduce the oracle (scores of rows holding that column shrink)."""
    from doppelspeller_b200 import encode, synthetic
    from oracle import oracle
    truth = synthetic.generate_truth_titles(12000, seed=61)
    test, _ = synthetic.generate_test_titles(truth, 300, seed=62)
    enc = encode.encode_canonical(test, truth)
    enc['idf64'] = enc['idf64'].copy()
    common = int(np.bincount(enc['t_cols'].astype(np.int64)).argmax())
    enc['idf64'][common] = -0.75
    rows, count, kth, _ = _match_maker(enc, 10)._index.topn(enc['q_ptr'], enc['q_cols'], 10, with_details=True)
    want_rows, want_count, want_kth = oracle.topn(oracle_index_from_encoded(enc), 10)
    assert np.array_equal(count, want_count)
    assert np.array_equal(rows, want_rows)
    assert np.array_equal(kth, want_kth)


def test_repeated_truth_titles_overflow_tiers():
    """A truth DB that repeats titles thousands of times: every copy ties at the top score, the 256 / 1,024-entry candidate
    buffers overflow, the 4,096-entry retry takes the 1,500-fold title and the dense fallback the 6,000-fold one."""
    from doppelspeller_b200 import encode, synthetic
    from oracle import oracle
    truth = synthetic.generate_truth_titles(24000, seed=81)
    rng = np.random.default_rng(82)
    for title, copies in (('acme holdings ltd', 1500), ('northern lights trading company', 6000)):
        for r in rng.choice(len(truth), copies, replace=False):
            truth[r] = title
    test, _ = synthetic.generate_test_titles(truth, 200, seed=83)
    test[:4] = ['acme holdings ltd', 'northern lights trading company', 'acme holding ltd', 'northern light trading company']
    enc = encode.encode_canonical(test, truth)
    for k in (10, 100):
        rows, count, kth, _ = _match_maker(enc, k)._index.topn(enc['q_ptr'], enc['q_cols'], k, with_details=True)
        want_rows, want_count, want_kth = oracle.topn(oracle_index_from_encoded(enc), k)
        assert np.array_equal(count, want_count)
        assert np.array_equal(rows, want_rows)
        assert np.array_equal(kth, want_kth)


def test_long_title_index_keeps_the_dense_scan():
    """Blocks of 2,048 long titles hold more than 65,535 postings (the 16-bit segment offsets of the posting form):
    such an index stays on the dense row scan; queries with more than 32 trigrams go through both forms elsewhere."""
    from doppelspeller_b200 import encode, synthetic
    from oracle import oracle
    truth = synthetic.generate_long_titles(5000, seed=71)
    test, _ = synthetic.generate_test_titles(truth, 120, seed=72)
    enc = encode.encode_canonical(test, truth)
    assert np.diff(enc['t_ptr'])[:2048].sum() > 65535
    rows, count, kth, _ = _match_maker(enc, 10)._index.topn(enc['q_ptr'], enc['q_cols'], 10, with_details=True)
    want_rows, want_count, want_kth = oracle.topn(oracle_index_from_encoded(enc), 10)
    assert np.array_equal(count, want_count)
    assert np.array_equal(rows, want_rows)
    assert np.array_equal(kth, want_kth)


def _tiny_case(truth_sets, query_sets, n_vocab):
    """Hand-built index: column ids given directly, idf from document frequencies."""
    import math
    n = len(truth_sets)
    df = np.zeros(n_vocab)
    for s in truth_sets:
        for c in s:
            df[c] += 1
    idf = np.array([math.log(n / d) if d > 0 else 0.0 for d in df])
    idf[df == 0] = idf.max()

    def csr(sets):
        ptr = np.zeros(len(sets) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in sets], out=ptr[1:])
        cols = np.array([c for s in sets for c in s], dtype=np.uint16)
        return ptr, cols
    t_ptr, t_cols = csr(truth_sets)
    q_ptr, q_cols = csr(query_sets)
    return dict(idf64=idf, t_ptr=t_ptr, t_cols=t_cols, q_ptr=q_ptr, q_cols=q_cols)


def _check_against_oracle(enc, k):
    from oracle import oracle
    rows, count, kth, flags = _match_maker(enc, k)._index.topn(enc['q_ptr'], enc['q_cols'], k, with_details=True)
    want_rows, want_count, want_kth = oracle.topn(oracle_index_from_encoded(enc), k)
    assert np.array_equal(count, want_count)
    assert np.array_equal(rows, want_rows)
    assert np.array_equal(kth, want_kth)
    return rows, count, flags


def test_edge_massive_ties_and_dropped_argmax():
    # 400 identical truth rows tie at the k-th place: more than the retained list can hold -> exact rescan,
    # and (like the reference) the k HIGHEST rows win even though an earlier row scores higher
    truth = [[0, 1, 2, 3]] + [[0, 1, 5 + i % 3] for i in range(30)] + [[0, 1, 9]] * 400 + [[20 + i, 21 + i] for i in range(50)]
    queries = [[0, 1, 2, 3], [0, 1, 9], [0, 1], [0, 1, 9, 30, 31], [40, 41, 42]]
    rows, count, flags = _check_against_oracle(_tiny_case(truth, queries, 80), 10)
    assert (flags & 1).any()


def test_edge_fewer_than_k_positives_returns_last_rows():
    truth = [[i, i + 1] for i in range(0, 200, 2)]
    queries = [[0, 1], [500], [4, 5, 8]]
    enc = _tiny_case(truth, queries, 600)
    rows, count, flags = _check_against_oracle(enc, 10)
    assert rows[1].tolist() == list(range(99, 89, -1))
    assert (flags & 2).all()


def test_edge_fewer_rows_than_k_raises_like_reference():
    truth = [[0, 1], [1, 2], [2, 3]]
    enc = _tiny_case(truth, [[0, 1]], 8)
    rows, count, flags = _check_against_oracle(enc, 5)
    assert count[0] == 3
    mm = _match_maker(enc, 5)
    with pytest.raises(Exception, match=r'top_matches.shape\[0\] != self.top_n'):
        mm.get_closest_matches(0)


def test_edge_empty_and_ragged_queries():
    truth = [[i % 7, 7 + i % 5, 12 + i % 3] for i in range(300)] + [list(range(40, 140))]
    queries = [[], [0], list(range(40, 140)), [0, 7, 12], list(range(0, 15))]
    _check_against_oracle(_tiny_case(truth, queries, 150), 10)
    _check_against_oracle(_tiny_case(truth, queries, 150), 100)


def test_edge_near_ties_inside_float_buffer():
    # many rows whose scores differ in the last bits: exercises the 1e-6 band and the retained-list flag
    rng = np.random.default_rng(11)
    truth = [sorted(rng.choice(60, size=int(rng.integers(3, 9)), replace=False).tolist()) for _ in range(4000)]
    queries = [sorted(rng.choice(60, size=int(rng.integers(2, 12)), replace=False).tolist()) for _ in range(200)]
    _check_against_oracle(_tiny_case(truth, queries, 60), 10)
    _check_against_oracle(_tiny_case(truth, queries, 60), 100)


def test_device_resident_inputs_match_host_inputs(golden_matchmaker):
    import torch
    g = golden_matchmaker
    mm = _match_maker(_golden_enc(g), 10)
    q_ptr = torch.as_tensor(g['q_ptr']).cuda()
    q_cols = torch.as_tensor(g['q_cols']).cuda()
    rows, count = mm._index.topn(q_ptr, q_cols, 10)
    assert rows.is_cuda
    assert np.array_equal(rows.cpu().numpy(), g['top10_rows'].astype(np.int64))


def test_sharded_phases_on_one_gpu_match_single_index(golden_matchmaker):
    """Truth split into 3 shards held on the same GPU: local -> stack -> merge -> rescan -> combine."""
    import torch
    from doppelspeller_b200 import sharded
    from doppelspeller_b200.index import TruthIndex, topn_merge
    g = golden_matchmaker
    n = int(g['t_ptr'].shape[0]) - 1
    for k, key in ((10, 'top10_rows'), (100, 'top100_rows')):
        offs = sharded.shard_offsets(n, 3)
        shards = []
        for r in range(3):
            ptr, cols = sharded.slice_truth_csr(g['t_ptr'], g['t_cols'], int(offs[r]), int(offs[r + 1]))
            shards.append(TruthIndex(ptr, cols, g['w64'], row_offset=int(offs[r]), n_total=n))
        local = [s.topn_local(g['q_ptr'], g['q_cols'], k) for s in shards]
        all_score = np.stack([l[0] for l in local])
        all_row = np.stack([l[1] for l in local])
        rows, count, kth, thr, flags = topn_merge(all_score, all_row, k, n, q_mx=local[0][2])
        flagged = np.nonzero(flags & 1)[0]
        if flagged.size:
            per_rows, per_count = [], []
            for s in shards:
                lr = np.full((len(flags), k), -1, dtype=np.int64)
                lc = np.zeros(len(flags), dtype=np.int32)
                s.topn_rescan(g['q_ptr'], g['q_cols'], local[0][2], thr, flags, k, lr, lc)
                per_rows.append(lr[flagged])
                per_count.append(lc[flagged])
            fixed, fixed_count = sharded.combine_rescans(torch.as_tensor(np.stack(per_rows)),
                                                         torch.as_tensor(np.stack(per_count)), k)
            rows[flagged] = fixed.numpy()
            count[flagged] = fixed_count.numpy()
        assert np.array_equal(rows, g[key].astype(np.int64))
        assert (count == k).all()


# ------------------------------------------------------------------ K2 / K3: pair scoring
def test_indel_ratio_matches_reference(golden_pairs):
    from doppelspeller_b200 import feature_engineering as fe
    g = golden_pairs
    got, dist = fe.fast_levenshtein_ratio_batch(g['ratio_a'], g['ratio_b'], g['ratio_la'].astype(np.uint8),
                                                g['ratio_lb'].astype(np.uint8), with_distance=True)
    keep = g['ratio_la'] + g['ratio_lb'] > 0          # empty vs empty raises in the reference
    assert np.array_equal(got[keep], g['ratio_out'][keep])
    assert fe.fast_levenshtein_ratio(np.array([2] * 29 + [3] * 21, np.uint8), np.array([2] * 29 + [4] * 21, np.uint8)) == 58
    with pytest.raises(ZeroDivisionError):
        fe.fast_levenshtein_ratio(np.zeros(0, np.uint8), np.zeros(0, np.uint8))


def test_indel_ratio_random_against_oracle():
    from doppelspeller_b200 import feature_engineering as fe
    from oracle import oracle
    rng = np.random.default_rng(5)
    n = 20000
    la = rng.integers(0, 256, n).astype(np.uint8)
    lb = rng.integers(0, 256, n).astype(np.uint8)
    short = rng.random(n) < 0.7
    la[short] = rng.integers(1, 60, short.sum())
    lb[short] = rng.integers(1, 60, short.sum())
    alpha = rng.integers(2, 38, n)
    a = (rng.integers(0, 1 << 30, (n, 255)) % alpha[:, None]).astype(np.uint8)
    b = (rng.integers(0, 1 << 30, (n, 255)) % alpha[:, None]).astype(np.uint8)
    similar = rng.random(n) < 0.5
    b[similar] = a[similar]
    flips = rng.integers(0, 255, (n, 4))
    for j in range(4):
        b[np.arange(n), flips[:, j]] = rng.integers(0, 38, n)
    wild = rng.random(n) < 0.02                        # codes outside the 38-symbol alphabet -> literal DP path
    a[wild, 0] = 200
    got = fe.fast_levenshtein_ratio_batch(a, b, la, lb)
    want = oracle.indel_ratio_u8_batch(a, b, la, lb)
    assert np.array_equal(got, want)


def test_construct_features_matches_reference(golden_pairs):
    from doppelspeller_b200 import feature_engineering as fe
    g = golden_pairs
    out = np.zeros((len(g['feat_la']), 66), dtype=np.float32)
    ret = fe.construct_features(g['feat_la'], g['feat_lb'], g['feat_a'], g['feat_b'], g['feat_counts'], np.uint8(1),
                                np.uint32(g['feat_n_truth']), np.zeros(66, np.uint8), out)
    assert ret is out
    assert features_equal(out, g['feat_out'])


def test_construct_features_pairs_table_form(golden_pairs):
    from doppelspeller_b200 import feature_engineering as fe
    g = golden_pairs
    titles = [str(t) for t in g['feat_titles']]
    truths = [str(t) for t in g['feat_truths']]
    n = len(titles)
    idx = np.arange(n, dtype=np.int32)
    got = fe.construct_features_pairs(fe.encode_titles(titles), fe.encode_titles(truths), g['feat_counts'], idx, idx, 1,
                                      int(g['feat_n_truth']))
    assert features_equal(got, g['feat_out'])


def test_construct_features_device_resident_padded_layout(golden_pairs):
    import torch
    from doppelspeller_b200 import feature_engineering as fe
    g = golden_pairs
    dev = [torch.as_tensor(g[key]).cuda() for key in ('feat_la', 'feat_lb', 'feat_a', 'feat_b')]
    counts = torch.as_tensor(g['feat_counts'].view(np.int32)).cuda()
    got_dev = fe.construct_features(dev[0], dev[1], dev[2], dev[3], counts, 1, int(g['feat_n_truth']))
    assert features_equal(got_dev.cpu().numpy(), g['feat_out'])


def test_levenshtein_ratio_and_prematch_against_oracle(example_titles):
    from doppelspeller_b200 import common, predict
    from oracle import oracle
    rng = np.random.default_rng(9)
    truth, test = example_titles['truth_titles'], example_titles['test_titles']
    xs = [test[i] for i in rng.integers(0, len(test), 3000)]
    ys = [truth[i] for i in rng.integers(0, len(truth), 3000)]
    for i in range(0, 3000, 3):                         # near duplicates: the > 94 region and the token-sort branch
        ys[i] = xs[i][:-1] if i % 2 else ' '.join(reversed(xs[i].split()))
    xs += ['', 'abc', 'ab', 'abcdefgh', 'x\ty']
    ys += ['', 'abc', 'abcd', 'abcdefgx', 'x y']
    got = common.levenshtein_ratio_batch(xs, ys)
    want = np.array([oracle.levenshtein_ratio(x, y) for x, y in zip(xs, ys)])
    assert np.array_equal(got, want)
    assert common.levenshtein_ratio('coolblu bv', 'coolblue bv') == 95
    assert common.levenshtein_token_sort_ratio('bv coolblue', 'coolblue bv') == 100
    from tests.test_oracle_golden import PUBLISHED_LEVENSHTEIN_ANSWERS          # python-Levenshtein docstring, fuzzywuzzy README
    for x, y, ratio, token_sort in PUBLISHED_LEVENSHTEIN_ANSWERS:
        assert common.levenshtein_ratio(x, y) == ratio and common.levenshtein_ratio(y, x) == ratio
        assert token_sort is None or common.levenshtein_token_sort_ratio(x, y) == token_sort
    xs2, ys2 = xs[:-5], ys[:-5]
    got = predict.get_levenshtein_ratios(xs2, ys2)
    want = np.array([oracle.prematch_ratio(x, y) for x, y in zip(xs2, ys2)])
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ f1: GPU index encoding
def test_gpu_trigram_encoder_matches_host_encoder():
    from doppelspeller_b200 import encode, synthetic
    from doppelspeller_b200.index import TruthIndex
    truth = synthetic.generate_truth_titles(30000, seed=21) + synthetic.generate_long_titles(300, seed=22) + ['abc', 'aaaaaaa', 'ab ab ab ab']
    test, _ = synthetic.generate_test_titles(truth, 3000, seed=23)
    want = encode.encode_canonical(test, truth)
    got = encode.encode_canonical_device(test, truth)
    for key in ('t_ptr', 'q_ptr', 't_cols', 'q_cols', 'vocab_codes'):
        assert np.array_equal(got[key].cpu().numpy(), want[key]), key
    assert np.array_equal(got['idf64'].cpu().numpy().view(np.uint64), want['idf64'].view(np.uint64))     # same libm log as math.log
    # device-resident encoding -> device-resident index -> candidates, no host copies in between
    index = TruthIndex(got['t_ptr'], got['t_cols'], got['idf64'])
    rows, count = index.topn(got['q_ptr'], got['q_cols'], 10)
    host_index = TruthIndex(want['t_ptr'], want['t_cols'], want['idf64'])
    want_rows, want_count = host_index.topn(want['q_ptr'], want['q_cols'], 10)
    assert np.array_equal(rows.cpu().numpy(), want_rows)
    with pytest.raises(Exception, match='outside'):
        encode.encode_canonical_device(['bad\ttitle'], truth[:10])


def test_large_vocabulary_uses_the_512_thread_scan():
    """n_vocab near the 50,653 trigram maximum: the slot table alone is > 100 KB, so the scan runs as one
    512-thread CTA per SM with the opt-in shared-memory carve-out."""
    rng = np.random.default_rng(13)
    n_vocab = 50600
    hot = rng.choice(n_vocab, size=300, replace=False)           # a Zipf-like head so that scores collide and rank
    def row(n):
        cols = set(rng.choice(hot, size=n // 2, replace=False).tolist()) | set(rng.integers(0, n_vocab, size=n - n // 2).tolist())
        return sorted(cols)
    truth = [row(int(rng.integers(4, 30))) for _ in range(6000)]
    queries = [row(int(rng.integers(3, 40))) for _ in range(150)] + [truth[i] for i in (5, 77, 4000)]
    _check_against_oracle(_tiny_case(truth, queries, n_vocab), 10)
    _check_against_oracle(_tiny_case(truth, queries, n_vocab), 100)


def test_edge_degenerate_weights_and_empty_rows():
    """All-zero idf (every trigram in every truth title: scores are 0/0 = NaN, nothing qualifies and the reference
    raises), a single truth row, empty truth rows, duplicated truth rows."""
    from oracle import oracle
    # identical truth titles -> idf = log(N / N) = 0 everywhere -> NaN scores -> no row passes `array >= threshold`
    same = _tiny_case([[0, 1, 2]] * 5, [[0, 1], [2], []], 3)
    assert (same['idf64'] == 0).all()
    rows, count, kth, flags = _match_maker(same, 2)._index.topn(same['q_ptr'], same['q_cols'], 2, with_details=True)
    want_rows, want_count, _ = oracle.topn(oracle_index_from_encoded(same), 2)
    assert np.array_equal(count, want_count) and (count == 0).all()
    with pytest.raises(Exception, match=r'top_matches.shape\[0\] != self.top_n'):
        _match_maker(same, 2).get_closest_matches(0)
    # one truth row
    _check_against_oracle(_tiny_case([[0, 1, 2]], [[0, 1], [5]], 8), 1)
    # empty truth rows between real ones, duplicates, a query equal to a truth row
    truth = [[], [0, 1, 2], [], [3, 4], [0, 1, 2], [], [5], [0, 1, 2], []] * 30
    queries = [[0, 1, 2], [3], [], [5, 6], [0, 4, 5]]
    for k in (1, 3, 10, 100):
        _check_against_oracle(_tiny_case(truth, queries, 8), k)


def test_candidate_buffer_overflow_falls_back_to_bounded_mode():
    """5,000 identical truth rows late in the DB tie at the top of a query: the per-launch candidate buffer and
    the rescan buffer both overflow, so the bounded-memory (dense, fixed-size chunk) passes must take over."""
    rng = np.random.default_rng(17)
    n_vocab = 400
    truth = [sorted(rng.choice(n_vocab, size=int(rng.integers(3, 12)), replace=False).tolist()) for _ in range(20000)]
    hot = [5, 17, 33, 90, 200, 201]
    for r in range(9000, 14000):
        truth[r] = hot
    queries = [hot, hot[:4], sorted(rng.choice(n_vocab, size=8, replace=False).tolist()), [5, 17, 350]]
    for k in (10, 100):
        rows, count, flags = _check_against_oracle(_tiny_case(truth, queries, n_vocab), k)
        assert rows[0].tolist() == list(range(13999, 13999 - k, -1))
        assert flags[0] & 1


def test_randomised_shapes_against_oracle():
    """Fuzz over DB size, batch size, top_n, vocabulary and duplication."""
    rng = np.random.default_rng(23)
    for trial in range(24):
        n_truth = int(rng.choice([1, 2, 7, 40, 300, 1500, 6000]))
        n_vocab = int(rng.choice([3, 12, 60, 500, 5000]))
        n_q = int(rng.integers(1, 120))
        k = int(rng.choice([1, 2, 5, 10, 37, 100, 250]))
        def row(max_len):
            n = int(rng.integers(0, min(max_len, n_vocab) + 1))
            return sorted(rng.choice(n_vocab, size=n, replace=False).tolist())
        truth = [row(30) for _ in range(n_truth)]
        if trial % 3 == 0 and n_truth > 4:                      # duplicates
            for r in rng.integers(0, n_truth, n_truth // 2):
                truth[int(r)] = truth[0]
        if not any(truth):
            truth[0] = [0]
        queries = [row(60) for _ in range(n_q)]
        _check_against_oracle(_tiny_case(truth, queries, n_vocab), k)


def test_randomised_posting_shapes_against_oracle():
    """Fuzz of the posting-list path (k_post needs more than 8,192 rows): vocabularies from 12 columns (every column dense:
    the accumulators stay zero and the dense patterns decide alone) to 40,000 uniformly drawn ones (no dense column at
    all), skewed vocabularies in between, queries of up to 250 columns (more than one 32-column group per block), duplicated
    rows (ties, overflowing candidate buffers), top_n from 1 to 250."""
    rng = np.random.default_rng(29)
    for trial in range(14):
        n_truth = int(rng.choice([8200, 9000, 12345, 20000]))
        n_vocab = int(rng.choice([12, 60, 500, 5000, 40000]))
        n_q = int(rng.integers(8, 160))
        k = int(rng.choice([1, 5, 10, 37, 100, 250]))
        skewed = trial % 2 == 1
        weights = 1.0 / np.arange(1, n_vocab + 1) ** (1.1 if skewed else 0.0)
        weights /= weights.sum()

        def rows_of(count, max_len):
            lengths = rng.integers(0, min(max_len, n_vocab) + 1, count)
            drawn = rng.choice(n_vocab, size=int(lengths.sum()), p=weights)
            return [sorted(set(part.tolist())) for part in np.split(drawn, np.cumsum(lengths)[:-1])]
        truth = rows_of(n_truth, 30 if trial % 4 else 120)
        if trial % 3 == 0:                                          # duplicates: hundreds of rows tie at the top
            for r in rng.integers(0, n_truth, n_truth // 3):
                truth[int(r)] = truth[int(r) % 7]
        if not any(truth):
            truth[0] = [0]
        queries = rows_of(n_q, 250 if trial % 5 == 0 else 40)
        queries[0] = truth[5]                                       # an exact copy of a truth row
        _check_against_oracle(_tiny_case(truth, queries, n_vocab), k)


def test_reference_idf_word_vector_on_the_gpu():
    """idf_word('first') = log(3 / 2) = 0.40547 (doppelspeller/tests/test_common.py:25-28) as construct_features emits it."""
    import math
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200.pipeline import truth_word_counts
    truth = ['first second first third first', 'first first', 'fifth']
    counts = truth_word_counts(truth)
    title = 'first second'
    la, lb = np.array([len(title)], np.uint8), np.array([len(truth[0])], np.uint8)
    feats = fe.construct_features(la, lb, fe.encode_title(title)[None], fe.encode_title(truth[0])[None], counts[:1], fe.SPACE_CODE, len(truth))
    idf = np.asarray(feats)[0, 6 + 2 * 15:6 + 3 * 15]
    assert round(float(idf[0]), 5) == 0.40547
    assert np.isclose(idf[0], math.log(3 / 2), rtol=1e-6) and np.isclose(idf[1], math.log(3.0), rtol=1e-6) and np.isnan(idf[5:]).all()


# ------------------------------------------------------------------ f3: transform_title
def test_transform_titles_match_reference(golden_transform):
    """ds_transform_titles (k_transform) against the reference's transform_title outputs: host tables."""
    from doppelspeller_b200 import common
    titles, outputs = golden_transform
    assert common.transform_titles(titles) == outputs
    assert common.transform_title(titles[4500]) == outputs[4500]
    assert common.transform_titles([]) == []


def test_transform_titles_device_table_matches_reference(golden_transform):
    """the same with the output table left on the device (the form the encoder and the pair kernels consume)"""
    import torch
    from doppelspeller_b200 import common
    titles, outputs = golden_transform
    out, off, raw = common.transform_titles_table(titles, device=torch.cuda.current_device(), warn=False)
    out, off = out.cpu().numpy(), off.cpu().numpy()
    blob = out.tobytes().decode('latin-1')
    assert [blob[off[i]:off[i + 1]] for i in range(len(titles))] == outputs
    # raw_len = len(text) before the [:255] cut (drives the reference's two warnings)
    assert int(raw[titles.index('x' * 256)]) == 256 and int(raw[titles.index('--')]) == 0


# ------------------------------------------------------------------ f4: boosted-tree inference
def _random_forest(rng, n_trees, n_features, max_depth):
    trees = []
    for _ in range(n_trees):
        tree, frontier = [], [(0, 0)]
        nodes = {0: None}
        next_id = 1
        while frontier:
            node, depth = frontier.pop(0)
            if depth < max_depth and rng.random() < 0.8:
                yes, no = next_id, next_id + 1
                next_id += 2
                nodes[node] = (int(rng.integers(n_features)), float(np.float32(rng.normal(0, 40))), yes, no, yes if rng.random() < 0.5 else no)
                frontier += [(yes, depth + 1), (no, depth + 1)]
            else:
                nodes[node] = (-1, float(np.float32(rng.normal(0, 0.3))), 0, 0, 0)
        for i in range(next_id):
            tree.append(nodes[i])
        trees.append(tree)
    return trees


@pytest.mark.parametrize('n_trees,max_depth', [(300, 5), (1000, 5), (3, 12), (0, 1)])
def _gbdt_case(n_trees, max_depth):
    from doppelspeller_b200 import gbdt
    rng = np.random.default_rng(100 + n_trees)
    model = gbdt.GbdtModel.from_trees(_random_forest(rng, n_trees, 66, max_depth), base_margin=-0.4, transform=gbdt.LOGISTIC)
    x = rng.normal(0, 50, size=(20011, 66)).astype(np.float32)
    x[rng.random(x.shape) < 0.08] = np.nan
    x[rng.random(x.shape) < 0.01] = np.inf
    x[:, 5] = np.round(x[:, 5])                      # values that land exactly on thresholds exercise the strict `<`
    return model, x


@pytest.mark.parametrize('n_trees,max_depth', [(300, 5), (3, 12)])
def test_gbdt_predict_device_resident_features(n_trees, max_depth):
    import torch
    model, x = _gbdt_case(n_trees, max_depth)
    got_dev = model.predict(torch.as_tensor(x).cuda()).cpu().numpy()
    assert np.array_equal(got_dev.view(np.uint32), model.predict(x).view(np.uint32))


@pytest.mark.parametrize('n_trees,max_depth', [(300, 5), (1000, 5), (3, 12), (0, 1)])
def test_gbdt_predict_matches_oracle(n_trees, max_depth):
    from doppelspeller_b200 import gbdt
    from oracle import oracle
    model, x = _gbdt_case(n_trees, max_depth)
    want = oracle.gbdt_predict(x, model.nodes, model.tree_offsets, model.base_margin, logistic=True)
    got = model.predict(x)
    assert np.allclose(got, want, rtol=1e-6, atol=0)
    margin_model = gbdt.GbdtModel(model.nodes, model.tree_offsets, model.base_margin, gbdt.MARGIN)
    want_margin = oracle.gbdt_predict(x, model.nodes, model.tree_offsets, model.base_margin, logistic=False)
    assert np.array_equal(margin_model.predict(x).view(np.uint32), want_margin.view(np.uint32))   # float32 sums: bit exact

