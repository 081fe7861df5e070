"""CPU, build container only: the oracle restatement against the UNMODIFIED reference imported from
/root/reference (numba kernels executed here).  Skipped where the reference is absent (GPU box) - there
the golden vectors minted from these same functions pin the oracle (test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle import oracle, ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason='/root/reference is not present')

N_QUERIES = 200


@pytest.fixture(scope='module')
def ref():
    return ref_import.import_reference()


@pytest.fixture(scope='module')
def example(ref):
    c = ref.constants
    truth = ref.common.get_ground_truth()
    test = ref.common.get_test_data().iloc[:N_QUERIES].copy()
    mm = ref.match_maker.MatchMaker(test.copy(), truth.copy(), 100)
    enc = oracle.encode_reference_order(list(test[c.COLUMN_N_GRAMS]), list(truth[c.COLUMN_N_GRAMS]))
    return dict(truth=truth, test=test, mm=mm, index=oracle.finish_index(enc), enc=enc)


def test_index_quantities(example):
    mm, index, enc = example['mm'], example['index'], example['enc']
    assert [mm.n_grams_decoding[i] for i in range(len(enc['vocab']))] == enc['vocab']          # match_maker.py:144-147
    assert np.array_equal(mm.sums_matrix_truth.view(np.uint32), index['sums'].view(np.uint32))  # :172-174
    for q in range(N_QUERIES):
        assert np.array_equal(mm.matrix_non_zero_columns[q], index['qs_cols'][index['qs_ptr'][q]:index['qs_ptr'][q + 1]])
    for v in range(0, len(enc['vocab']), 7):
        rows = mm.matrix_truth_non_zero_columns_and_values[v][0]
        assert np.array_equal(rows, index['post_rows'][index['post_ptr'][v]:index['post_ptr'][v + 1]])


def test_python_sum_restatement(example):
    mm, index = example['mm'], example['index']
    for q in range(N_QUERIES):
        nz = mm.matrix_non_zero_columns[q]
        want = sum([mm._get_idf_given_index(r) for r in nz])                                   # match_maker.py:197
        assert oracle.py_float_sum(index['w64'], nz) == want


def test_fast_jaccard_bits(ref, example):
    mm, index = example['mm'], example['index']
    for q in range(0, N_QUERIES, 20):
        nz = mm.matrix_non_zero_columns[q]
        mx = sum([mm._get_idf_given_index(r) for r in nz])
        want = ref.match_maker.fast_jaccard(mm.number_of_truth_titles, mx, nz, mm.matrix_truth_non_zero_columns_and_values,
                                            mm.sums_matrix_truth)
        got = oracle.fast_jaccard(index, q)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize('k', [10, 100])
def test_candidate_lists(ref, example, k):
    mm, index, truth = example['mm'], example['index'], example['truth']
    title_ids = truth[ref.constants.COLUMN_TITLE_ID].to_numpy()
    rows, count, _ = oracle.topn(index, k)
    mm.top_n = k
    for q in range(N_QUERIES):
        assert mm.get_closest_matches(q) == title_ids[rows[q]].tolist()


@pytest.mark.parametrize('n_truth,n_q,k,seed', [(3000, 120, 10, 41), (900, 60, 100, 42), (64, 40, 64, 43)])
def test_candidate_lists_on_the_synthetic_workload(ref, n_truth, n_q, k, seed):
    """The reference's own MatchMaker on titles of the BENCH generator (doppelspeller_b200.synthetic: heavier duplication,
    pseudo-words, edited copies of truth titles as queries - shapes the example data does not have), row by row against
    the oracle; the last case has top_n == number of truth rows (every row returned, match_maker.py:66-71)."""
    import pandas as pd

    from doppelspeller_b200 import synthetic
    c = ref.constants
    truth_titles = synthetic.generate_truth_titles(n_truth, seed=seed)
    test_titles, _ = synthetic.generate_test_titles(truth_titles, n_q, seed=seed + 100)

    def frame(titles, first_id):
        return pd.DataFrame({c.COLUMN_TITLE_ID: np.arange(first_id, first_id + len(titles)),
                             c.COLUMN_TRANSFORMED_TITLE: list(titles),
                             c.COLUMN_N_GRAMS: [ref.common.get_n_grams(t, 3) for t in titles]})
    truth, test = frame(truth_titles, 5000), frame(test_titles, 0)
    mm = ref.match_maker.MatchMaker(test.copy(), truth.copy(), k)
    index = oracle.finish_index(oracle.encode_reference_order(list(test[c.COLUMN_N_GRAMS]), list(truth[c.COLUMN_N_GRAMS])))
    assert np.array_equal(mm.sums_matrix_truth.view(np.uint32), index['sums'].view(np.uint32))
    rows, count, _ = oracle.topn(index, k)
    assert (count == k).all()
    title_ids = truth[c.COLUMN_TITLE_ID].to_numpy()
    for q in range(n_q):
        assert mm.get_closest_matches(q) == title_ids[rows[q]].tolist()


def test_fast_arg_top_k_random(ref):
    rng = np.random.default_rng(0)
    fatk = ref.match_maker.fast_arg_top_k
    for trial in range(200):
        n, k = int(rng.integers(1, 400)), int(rng.integers(1, 50))
        v = rng.random(n)
        if trial % 3 == 0:
            v = np.round(v, 1)
        if trial % 5 == 0:
            v[rng.random(n) < 0.8] = 0.0
        if trial % 7 == 0:
            v = 0.25 + rng.integers(-4, 5, n) * 3e-7
        assert np.array_equal(oracle.fast_arg_top_k(v, k), fatk(v, k))


def test_fast_levenshtein_ratio_random(ref):
    rng = np.random.default_rng(1)
    flr = ref.feature_engineering.fast_levenshtein_ratio
    for trial in range(3000):
        big = trial % 5 == 0
        la, lb = int(rng.integers(1, 256 if big else 40)), int(rng.integers(1, 256 if big else 40))
        alpha = int(rng.integers(2, 8))
        a = rng.integers(1, 1 + alpha, la).astype(np.uint8)
        b = rng.integers(1, 1 + alpha, lb).astype(np.uint8)
        assert oracle.indel_ratio_u8(a, b) == int(flr(a, b))


def test_levenshtein_ratio_distance_recovered_from_the_reference_kernel(ref):
    """a5 cross-check of SURVEY 8(c): python-levenshtein is absent, but the reference's own jitted fast_levenshtein_ratio
    computes the same distance d (feature_engineering.py:50-61).  For la + lb <= 100 consecutive admissible d (same parity
    as la + lb) move the truncated ratio by >= 2, so d is recovered uniquely from the reference's output; common.py:162's
    expression int(round(ratio * 100)) with ratio = (la + lb - d) / (la + lb), evaluated by Python itself, must then be
    what oracle.levenshtein_ratio returns (distance AND half-even rounding)."""
    rng = np.random.default_rng(5)
    flr = ref.feature_engineering.fast_levenshtein_ratio
    alphabet = np.frombuffer(b'abcdefgh ', dtype=np.uint8)
    halves = 0
    for trial in range(4000):
        la = int(rng.integers(1, 50))
        lb = int(rng.integers(1, 101 - la)) if trial % 4 else la            # equal lengths: the x.5 ties (e.g. 14/16)
        a = alphabet[rng.integers(0, int(rng.integers(2, 10)), la)]
        b = a.copy()[:lb] if trial % 3 == 0 and lb <= la else alphabet[rng.integers(0, 9, lb)]
        if trial % 3 == 0:
            b = b.copy()
            b[rng.integers(0, len(b), 1 + len(b) // 8)] = alphabet[int(rng.integers(0, 9))]
        total = la + len(b)
        got = int(flr(np.ascontiguousarray(a), np.ascontiguousarray(b)))
        candidates = [d for d in range(total % 2, total + 1, 2) if -1e-9 <= 100.0 * (total - d) / total - got < 1.0 + 1e-9]
        assert len(candidates) == 1, (total, got, candidates)
        want = int(round((total - candidates[0]) / total * 100))
        halves += (200 * (total - candidates[0])) % (2 * total) == total
        assert oracle.levenshtein_ratio(a.tobytes().decode(), b.tobytes().decode()) == want
    assert halves > 20                                                       # the half-even branch was exercised


def test_construct_features_random(ref, example):
    rng = np.random.default_rng(2)
    c = ref.constants
    truth_titles = list(example['truth'][c.COLUMN_TRANSFORMED_TITLE])
    test_titles = list(example['test'][c.COLUMN_TRANSFORMED_TITLE])
    pairs = [(test_titles[i % N_QUERIES], truth_titles[int(rng.integers(0, len(truth_titles)))]) for i in range(600)]
    words = ['ab', 'cde', 'fghi', 'jk', 'lmnop', 'q', 'rst', 'uv', 'wxyz', 'a1', 'b22', 'c333', 'dd', 'ee', 'ffg', 'hh']
    pairs += [(' '.join(rng.choice(words, max(1, n - 1))), ' '.join(rng.choice(words, n))) for n in range(1, 21) for _ in range(5)]
    # C4's stress slice and beyond: 65..255 characters, incl. la + lb > 255 (the uint8 cells of the DP wrap) and > 15 words
    for n_a, n_b in [(int(rng.integers(14, 70)), int(rng.integers(14, 70))) for _ in range(160)]:
        a, b = ' '.join(rng.choice(words, n_a))[:255].strip(), ' '.join(rng.choice(words, n_b))[:255].strip()
        pairs.append((a, b if rng.random() < 0.5 else (a[:int(rng.integers(1, len(a)))].strip() + ' ' + b[:100].strip())[:255].strip()))
    assert sum(len(a) + len(b) > 255 for a, b in pairs) > 20 and max(len(b) for a, b in pairs) <= 255
    la = np.array([len(a) for a, b in pairs], np.uint8)
    lb = np.array([len(b) for a, b in pairs], np.uint8)
    ta = np.vstack([oracle.encode_title(a) for a, b in pairs])
    tb = np.vstack([oracle.encode_title(b) for a, b in pairs])
    counts = rng.integers(0, 3000, (len(pairs), 15)).astype(np.uint32)
    want = np.zeros((len(pairs), 66), np.float32)
    with np.errstate(all='ignore'):
        ref.feature_engineering.construct_features(la, lb, ta, tb, counts, np.uint8(1), np.uint32(30000),
                                                   np.zeros(66, np.uint8), want)
    got = oracle.construct_features(la, lb, ta, tb, counts, 1, 30000)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all()


def test_reference_golden_tests_still_hold(ref):
    # the reference's own three tests (doppelspeller/tests/test_common.py:16-28)
    title = '''LKJblksd skjasl dfkjf &* 8*&&&8 GGdjsdkj--sdsd-"sdi..//' d'  k   bkjh77_asda33'''
    assert ref.common.transform_title(title) == 'lkjblksd skjasl dfkjf 88 ggdjsdkj sdsd sdi d k bkjh77asda33'


def test_transform_title_random_unicode(ref):
    """oracle.transform_title against the reference's regex formulation on random code point soup."""
    import logging
    logging.disable(logging.CRITICAL)
    try:
        rng = np.random.default_rng(77)
        ranges = [(0, 0x80), (0x80, 0x250), (0x300, 0x370), (0x1E00, 0x2300), (0x3000, 0x3100), (0xFB00, 0xFB10), (0x1F600, 0x1F610)]
        for _ in range(1500):
            n = int(rng.integers(0, 80))
            chars = []
            for _ in range(n):
                lo, hi = ranges[int(rng.integers(len(ranges)))]
                chars.append(chr(int(rng.integers(lo, hi))))
            title = ''.join(chars)
            assert oracle.transform_title(title) == ref.common.transform_title(title), repr(title)
    finally:
        logging.disable(logging.NOTSET)


def test_prematch_cascade_against_reference_prediction(ref, example):
    """a6: oracle.prematch_ratio against the reference's own Prediction._get_levenshtein_ratio /
    _get_levenshtein_deletion_ratio (predict.py:140-156), and the repo's host selection against the reference's
    `> 94` / group-max / ambiguity drop (predict.py:158-176, run through pandas exactly as written there).
    python-levenshtein's `ratio` itself stays a shim (third party, absent): what is pinned here is the cascade -
    the float64 length pre-filter, the `<= 94` switch to the token-sorted ratio and the selection."""
    import pandas as pd
    from doppelspeller_b200 import predict as ours
    c, s = ref.constants, ref.settings
    prediction = ref.predict.Prediction
    rng = np.random.default_rng(5)
    truth_titles = list(example['truth'][c.COLUMN_TRANSFORMED_TITLE])
    test_titles = list(example['test'][c.COLUMN_TRANSFORMED_TITLE])
    rows, _, _ = oracle.topn(example['index'], 100)
    pairs = [(test_titles[q], truth_titles[r]) for q in range(60) for r in rows[q]]
    # near-identical pairs exercise the `ratio > 94` and the token-sort branches, word swaps only the latter
    for title in truth_titles[:400]:
        words = title.split()
        pairs.append((title, title[:-1]))
        pairs.append((title.replace('e', 'a', 1), title))
        if len(words) > 1:
            pairs.append((' '.join(words[::-1]), title))
            pairs.append((' '.join(words[1:] + words[:1]) + 'x', title))
    for _ in range(400):                                     # lengths around the pre-filter boundary (predict.py:150)
        la = int(rng.integers(1, 120))
        lb = max(1, int(round(la * rng.uniform(0.85, 1.0))))
        pairs.append(('a' * la, 'a' * lb))
    got = np.array([oracle.prematch_ratio(x, y) for x, y in pairs])
    want = np.array([prediction._get_levenshtein_ratio(x, y) for x, y in pairs])
    assert np.array_equal(got, want)
    for x, y in pairs[::97]:
        want_deletion = prediction._get_levenshtein_deletion_ratio(x, y)
        assert ours.get_levenshtein_deletion_ratios([len(x)], [len(y)])[0] == want_deletion
    # selection: the reference's own pandas statements on a frame of (test_index, ratio)
    test_index = np.concatenate([np.repeat(np.arange(60), 100), 60 + rng.integers(0, 300, len(pairs) - 6000)])
    frame = pd.DataFrame({c.COLUMN_TEST_INDEX: test_index, c.COLUMN_LEVENSHTEIN_RATIO: want})
    matches = frame.loc[frame[c.COLUMN_LEVENSHTEIN_RATIO] > s.LEVENSHTEIN_RATIO_THRESHOLD, :]          # predict.py:172
    is_max = matches.groupby([c.COLUMN_TEST_INDEX])[c.COLUMN_LEVENSHTEIN_RATIO].transform('max') == \
        matches[c.COLUMN_LEVENSHTEIN_RATIO]                                                             # :173-174
    kept = prediction._remove_duplicated_matches(matches.loc[is_max, :])                               # :176
    assert np.array_equal(ours.select_close_matches(test_index, want), np.sort(kept.index.to_numpy()))
