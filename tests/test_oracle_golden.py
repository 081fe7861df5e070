"""CPU: the oracle restatement against the golden vectors minted from the reference
(tests/golden/make_golden.py) - this is what pins the oracle on the GPU box, where /root/reference
does not exist."""
import numpy as np

from oracle import oracle
from tests.conftest import features_equal, oracle_index_from_golden


def test_sums_and_mx_match_reference(golden_matchmaker):
    g = golden_matchmaker
    index = oracle_index_from_golden(g)
    assert np.array_equal(index['sums'].view(np.uint32), g['sums'].view(np.uint32))      # match_maker.py:172-174
    mx = np.array([oracle.py_float_sum(g['w64'], index['qs_cols'][index['qs_ptr'][q]:index['qs_ptr'][q + 1]])
                   for q in range(len(g['mx']))])
    assert np.array_equal(mx.view(np.uint64), g['mx'].view(np.uint64))                    # match_maker.py:197


def test_fast_jaccard_bit_exact(golden_matchmaker):
    g = golden_matchmaker
    index = oracle_index_from_golden(g)
    for i, q in enumerate(g['jac_queries']):
        got = oracle.fast_jaccard(index, int(q))
        assert np.array_equal(got.view(np.uint64), g['jac'][i].view(np.uint64))


def test_candidate_lists_match_reference(golden_matchmaker):
    g = golden_matchmaker
    index = oracle_index_from_golden(g)
    for k, key in ((10, 'top10_rows'), (100, 'top100_rows')):
        rows, count, _ = oracle.topn(index, k)
        assert (count == k).all()
        assert np.array_equal(rows, g[key].astype(np.int64))


def test_fast_arg_top_k_known_answers(golden_topk):
    g = golden_topk
    for i in range(int(g['n_cases'])):
        got = oracle.fast_arg_top_k(g[f'v{i}'], int(g[f'k{i}']))
        assert np.array_equal(got, g[f'r{i}']), f'case {i}'
    # the hand-checked vectors of SURVEY.md 8(c)
    assert oracle.fast_arg_top_k(np.array([0.9, 0.5, 0.5, 0.5]), 2).tolist() == [3, 2]
    assert oracle.fast_arg_top_k(np.zeros(8), 3).tolist() == [7, 6, 5]
    assert oracle.fast_arg_top_k(np.array([0, 0.3, 0, 0.2, 0, 0]), 3).tolist() == [5, 4, 3]
    assert oracle.fast_arg_top_k(np.array([0.5, 0.5000004, 0.4999996, 0.1, 0.9]), 2).tolist() == [4, 2]
    assert oracle.fast_arg_top_k(np.array([0.1, 0.2]), 3).tolist() == [1, 0]


def test_indel_ratio_matches_reference(golden_pairs):
    g = golden_pairs
    for i in range(len(g['ratio_out'])):
        a = g['ratio_a'][i, :g['ratio_la'][i]]
        b = g['ratio_b'][i, :g['ratio_lb'][i]]
        assert oracle.indel_ratio_u8(a, b) == int(g['ratio_out'][i]), f'pair {i}'
    # SURVEY.md 8(c) wrap vectors
    u8 = lambda v, n: np.full(n, v, dtype=np.uint8)   # noqa: E731
    assert oracle.indel_ratio_u8(u8(2, 128), u8(3, 128)) == 100
    assert oracle.indel_ratio_u8(u8(2, 127), u8(3, 128)) == 0
    assert oracle.indel_ratio_u8(u8(2, 200), u8(3, 100)) == 85
    assert oracle.indel_ratio_u8(u8(2, 255), u8(3, 255)) == 50


def test_construct_features_matches_reference(golden_pairs):
    g = golden_pairs
    got = oracle.construct_features(g['feat_la'], g['feat_lb'], g['feat_a'], g['feat_b'], g['feat_counts'], 1,
                                    int(g['feat_n_truth']))
    assert features_equal(got, g['feat_out'])
    # the oracle is in fact bit-identical to the reference on this host (same libm log)
    same = (got.view(np.uint32) == g['feat_out'].view(np.uint32)) | (np.isnan(got) & np.isnan(g['feat_out']))
    assert same.mean() > 0.9999


def test_docstring_encoding_vector():
    # feature_engineering.py:28-29: "coolblue bv" -> [4, 16, 16, 13, 3, 13, 22, 6, 1, 3, 23]
    assert oracle.encode_title('coolblue bv')[:11].tolist() == [4, 16, 16, 13, 3, 13, 22, 6, 1, 3, 23]
    assert oracle.encode_title('coolblue bv')[11:].sum() == 0


def test_levenshtein_ratio_known_answers():
    # python-levenshtein semantics (parity unpinned: restated, see oracle/ds_oracle.c)
    assert oracle.levenshtein_ratio('', '') == 100
    assert oracle.levenshtein_ratio('abc', 'abc') == 100
    assert oracle.levenshtein_ratio('abc', 'xyz') == 0
    assert oracle.levenshtein_ratio('coolblu bv', 'coolblue bv') == 95          # 20/21 = 95.24
    assert oracle.levenshtein_ratio('ab', 'abcd') == 67                          # 4/6 = 66.67
    assert oracle.levenshtein_ratio('abcdefgh', 'abcdefgx') == 88                # 14/16 = 87.5 -> half-even 88
    assert oracle.levenshtein_ratio('abcd', 'abcx') == 75
    assert oracle.levenshtein_token_sort_ratio('bv coolblue', 'coolblue bv') == 100
    assert oracle.prematch_ratio('a' * 10, 'a' * 30) == 0                        # length pre-filter (predict.py:150)


# The only answers the absent third-party code itself publishes: python-Levenshtein's docstring of ratio()
# (0.583333... and 0.0) and fuzzywuzzy's README, whose fuzz.ratio / token_sort_ratio are int(round(100 * Levenshtein ratio))
# when python-Levenshtein is installed - the same expression as common.py:161-167.  Anchors, not a pin: see DESIGN.md section 2.
PUBLISHED_LEVENSHTEIN_ANSWERS = (
    ('Hello world!', 'Holly grail!', 58, None), ('Brian', 'Jesus', 0, None),
    ('this is a test', 'this is a test!', 97, None),
    ('fuzzy wuzzy was a bear', 'wuzzy fuzzy was a bear', 91, 100),
)


def test_levenshtein_ratio_published_answers_of_the_third_party_libraries():
    for x, y, ratio, token_sort in PUBLISHED_LEVENSHTEIN_ANSWERS:
        assert oracle.levenshtein_ratio(x, y) == ratio
        assert oracle.levenshtein_ratio(y, x) == ratio
        if token_sort is not None:
            assert oracle.levenshtein_token_sort_ratio(x, y) == token_sort


def test_transform_title_matches_reference(golden_transform):
    """common.py:20-47 on 7,016 raw titles (example data + seeded accents / white space / length edge cases) and the
    reference's own test vector (doppelspeller/tests/test_common.py:16-19)."""
    titles, outputs = golden_transform
    assert [oracle.transform_title(t) for t in titles] == outputs
    assert oracle.transform_title('''LKJblksd skjasl dfkjf &* 8*&&&8 GGdjsdkj--sdsd-"sdi..//' d'  k   bkjh77_asda33''') == \
        'lkjblksd skjasl dfkjf 88 ggdjsdkj sdsd sdi d k bkjh77asda33'


def test_reference_word_counter_and_idf_word_vectors():
    """The reference's own known answers (doppelspeller/tests/test_common.py:21-28): document frequencies over per-title word
    SETS {'first': 2, 'second': 1, 'third': 1, 'fifth': 1} and idf_word('first') = log(3 / 2) = 0.40547 - here as they reach
    the hot path: the [n_truth, 15] count table of the pipeline and the idf features (columns 36..50) of construct_features."""
    import math
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200.pipeline import truth_word_counts
    truth = ['first second first third first', 'first first', 'fifth']
    counts = truth_word_counts(truth)
    assert counts[0, :5].tolist() == [2, 1, 2, 1, 2] and counts[1, :2].tolist() == [2, 2] and counts[2, 0] == 1
    assert not counts[0, 5:].any()
    title = 'first second'
    la, lb = np.array([len(title)], np.uint8), np.array([len(truth[0])], np.uint8)
    feats = oracle.construct_features(la, lb, fe.encode_title(title)[None], fe.encode_title(truth[0])[None], counts[:1], fe.SPACE_CODE, len(truth))
    idf = feats[0, 6 + 2 * 15:6 + 3 * 15]
    assert round(float(idf[0]), 5) == 0.40547 == round(math.log(3 / 2), 5)           # 'first'
    assert idf[0] == np.float32(math.log(3 / 2)) and idf[1] == np.float32(math.log(3 / 1))   # 'second'
    assert np.isnan(idf[5:]).all()
