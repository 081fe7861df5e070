"""CPU: host-side logic of the drop-in layer (no kernels): canonical trigram encoding, the length
pre-filter and the match selection of Prediction, title encoding, synthetic data determinism."""
import numpy as np
import pandas as pd

from oracle import oracle


def test_canonical_encoding_matches_get_n_grams():
    from doppelspeller_b200 import encode, synthetic
    truth = synthetic.generate_truth_titles(500, seed=5)
    test, _ = synthetic.generate_test_titles(truth, 80, seed=6)
    enc = encode.encode_canonical(test, truth)
    vocab = [encode.trigram_text(c) for c in enc['vocab_codes']]
    assert (np.diff(enc['vocab_codes']) > 0).all()          # column ids = ranks in code order (' ' < a..z < 0..9)
    for i in (0, 7, 311, 499):
        cols = enc['t_cols'][enc['t_ptr'][i]:enc['t_ptr'][i + 1]]
        assert {vocab[c] for c in cols} == oracle.get_n_grams(truth[i], 3)          # common.py:150-151
        assert (np.diff(cols.astype(np.int64)) > 0).all()
    df = np.bincount(enc['t_cols'], minlength=len(vocab))
    seen = df > 0
    want = np.array([np.log(len(truth) / d) for d in df[seen]])
    assert np.allclose(enc['idf64'][seen], want, rtol=0, atol=0) or np.array_equal(enc['idf64'][seen], want)
    assert (enc['idf64'][~seen] == enc['idf64'][seen].max()).all()                 # match_maker.py:151,180-181


def test_encode_title_docstring_vector():
    from doppelspeller_b200 import feature_engineering as fe
    assert fe.encode_title('coolblue bv')[:11].tolist() == [4, 16, 16, 13, 3, 13, 22, 6, 1, 3, 23]   # feature_engineering.py:28-29
    codes, offsets = fe.encode_titles(['coolblue bv', 'abc'])
    assert offsets.tolist() == [0, 11, 14] and codes[11:].tolist() == [2, 3, 4]
    assert fe.FEATURES_COUNT == 66 and fe.SPACE_CODE == 1
    assert fe.get_truth_words_counts('coolblue bv', {'coolblue': 1, 'bv': 2145})[:3].tolist() == [1, 2145, 0]


def test_length_prefilter_matches_reference_formula():
    from doppelspeller_b200 import predict
    rng = np.random.default_rng(0)
    la, lb = rng.integers(1, 256, 5000), rng.integers(1, 256, 5000)
    got = predict.get_levenshtein_deletion_ratios(la, lb) < 94
    want = np.array([((x + y - abs(x - y)) / (x + y)) * 100 < 94 for x, y in zip(la.tolist(), lb.tolist())])   # predict.py:140-151
    assert np.array_equal(got, want)
    assert all(bool(oracle.lib().orc_prefilter_rejects(int(x), int(y))) == w for x, y, w in zip(la[:500], lb[:500], want[:500]))


def test_select_close_matches_follows_pandas_logic():
    from doppelspeller_b200 import predict
    rng = np.random.default_rng(1)
    test_index = np.repeat(np.arange(200), 10)
    ratios = rng.choice([0, 90, 94, 95, 96, 97, 100], size=2000, p=[0.6, 0.1, 0.05, 0.1, 0.05, 0.05, 0.05])
    kept = predict.select_close_matches(test_index, ratios)
    frame = pd.DataFrame({'test_index': test_index, 'ratio': ratios})
    m = frame[frame['ratio'] > 94]
    m = m[m.groupby('test_index')['ratio'].transform('max') == m['ratio']]          # predict.py:172-174
    dup = m.loc[m.duplicated(['test_index']), 'test_index']                          # predict.py:158-161
    m = m[~m['test_index'].isin(dup)]
    assert np.array_equal(kept, m.index.to_numpy())


def test_synthetic_generator_is_deterministic_and_normalised():
    from doppelspeller_b200 import synthetic
    a = synthetic.generate_truth_titles(300, seed=9)
    assert a == synthetic.generate_truth_titles(300, seed=9)
    t1, s1 = synthetic.generate_test_titles(a, 100, seed=10)
    t2, s2 = synthetic.generate_test_titles(a, 100, seed=10)
    assert t1 == t2 and np.array_equal(s1, s2)
    for title in a + t1 + synthetic.generate_long_titles(20, seed=3):
        assert 3 <= len(title) <= 255 and title == ' '.join(title.split())
        assert set(title) <= set(' abcdefghijklmnopqrstuvwxyz0123456789')


def test_ascii_of_codepoint_table():
    """The table handed to ds_transform_titles: NFD + ascii-ignore per code point (common.py:25-26)."""
    import unicodedata
    from doppelspeller_b200 import common
    table = common.ascii_of_codepoint_table()
    assert table.dtype == np.uint8 and table.shape[0] == 0x2270
    assert all(table[c] == c for c in range(1, 128)) and table[0] == 0
    for ch, want in (('é', 'e'), ('Å', 'A'), ('ñ', 'n'), ('ø', ''), ('ß', ''), ('\u212a', 'K'), ('≠', '='), ('ﬁ', ''), ('株', '')):
        cp = ord(ch)
        got = chr(table[cp]) if cp < table.shape[0] and table[cp] else ''
        assert got == want == unicodedata.normalize('NFD', ch).encode('ascii', 'ignore').decode()


def test_gbdt_dump_parser_and_oracle_semantics():
    """xgboost JSON dump -> flat trees (children after their parent), and the restated prediction rule: strict `<`,
    NaN -> the `missing` child, float32 sum in tree order, sigmoid."""
    from doppelspeller_b200 import gbdt
    from oracle import oracle
    dump = ['{"nodeid": 0, "depth": 0, "split": "f2", "split_condition": 0.5, "yes": 1, "no": 2, "missing": 2, "children": ['
            '{"nodeid": 1, "leaf": -0.25}, {"nodeid": 2, "depth": 1, "split": "f0", "split_condition": 3, "yes": 3, "no": 4, '
            '"missing": 3, "children": [{"nodeid": 3, "leaf": 0.5}, {"nodeid": 4, "leaf": 1.5}]}]}',
            '{"nodeid": 0, "leaf": 0.125}']
    model = gbdt.GbdtModel.from_xgboost_dump(dump, base_score=0.5, objective='reg:logistic')
    assert model.n_trees == 2 and model.tree_offsets.tolist() == [0, 5, 6] and model.base_margin == 0.0
    assert model.nodes['feature'].tolist() == [2, -1, 0, -1, -1, -1]
    x = np.array([[0, 0, 0.25], [0, 0, 0.5], [3, 0, 0.5], [np.nan, 0, 0.75], [9, 0, np.nan], [np.nan, 0, np.nan]], dtype=np.float32)
    margins = oracle.gbdt_predict(x, model.nodes, model.tree_offsets, model.base_margin, logistic=False)
    assert margins.tolist() == [-0.125, 0.625, 1.625, 0.625, 1.625, 0.625]
    probabilities = oracle.gbdt_predict(x, model.nodes, model.tree_offsets, model.base_margin, logistic=True)
    assert np.allclose(probabilities, 1.0 / (1.0 + np.exp(-margins.astype(np.float64))), rtol=1e-6)


def test_model_match_selection():
    """predict.py:242-249: maximum of the test_index, above 0.9, attained once."""
    from doppelspeller_b200 import gbdt
    test_index = np.array([0, 0, 0, 1, 1, 2, 2, 3])
    predictions = np.array([0.95, 0.99, 0.2, 0.97, 0.97, 0.5, 0.89, 0.91], dtype=np.float32)
    assert gbdt.select_model_matches(test_index, predictions).tolist() == [1, 7]   # 1: tie at the maximum, 2: below 0.9


def test_gbdt_dump_feature_names_and_tree_limit():
    from doppelspeller_b200 import gbdt
    dump = ['{"nodeid": 0, "split": "ratio", "split_condition": 94.5, "yes": 1, "no": 2, "missing": 1, "children": ['
            '{"nodeid": 1, "leaf": -1.0}, {"nodeid": 2, "leaf": 2.0}]}', '{"nodeid": 0, "leaf": 7.0}']
    model = gbdt.GbdtModel.from_xgboost_dump(dump, base_score=0.25, objective='binary:logistic', ntree_limit=1,
                                             feature_names=['length', 'ratio'])
    assert model.n_trees == 1 and model.nodes['feature'].tolist() == [1, -1, -1] and model.transform == gbdt.LOGISTIC
    assert abs(model.base_margin - np.log(0.25 / 0.75)) < 1e-6
    linear = gbdt.GbdtModel.from_xgboost_dump(dump, base_score=0.5, objective='reg:squarederror', feature_names=['length', 'ratio'])
    assert linear.transform == gbdt.MARGIN and linear.base_margin == 0.5 and linear.n_trees == 2


def test_gbdt_model_rejects_malformed_trees():
    """The kernel walks device copies of the trees unchecked: GbdtModel validates them once on the host."""
    import pytest
    from doppelspeller_b200.gbdt import GbdtModel
    good = [[(3, 0.5, 1, 2, 1), (-1, 0.1, 0, 0, 0), (-1, -0.2, 0, 0, 0)]]
    model = GbdtModel.from_trees(good)
    assert model.n_trees == 1 and model.n_features_needed == 4
    with pytest.raises(ValueError):
        GbdtModel.from_trees([[(3, 0.5, 1, 5, 1), (-1, 0.1, 0, 0, 0), (-1, -0.2, 0, 0, 0)]])       # child outside the tree
    with pytest.raises(ValueError):
        GbdtModel.from_trees([[(3, 0.5, 0, 2, 1), (-1, 0.1, 0, 0, 0), (-1, -0.2, 0, 0, 0)]])       # child not after its parent
    with pytest.raises(ValueError):
        GbdtModel.from_trees([[]])                                                                  # empty tree
    with pytest.raises(ValueError):
        model.predict(np.zeros((2, 3), dtype=np.float32))                                           # feature 3 of a 3-column matrix


def test_combine_rescans_matches_the_plain_loop():
    """sharded.combine_rescans (vectorised, device side in production) against its definition: every shard reports its k
    highest qualifying rows in descending order; shards are ascending row ranges, so the answer is the concatenation from
    the highest shard down, cut at k, padded with -1."""
    import torch

    from doppelspeller_b200 import sharded
    rng = np.random.default_rng(17)
    for n_shards, n_f, k in ((1, 5, 3), (2, 40, 10), (3, 33, 1), (5, 64, 7), (8, 17, 100), (4, 0, 5)):
        counts = rng.integers(0, k + 1, size=(n_shards, n_f)).astype(np.int32)
        counts[:, : n_f // 4] = 0                                    # queries no shard has a row for
        counts[-1, n_f // 4: n_f // 2] = k                           # the highest shard alone fills the answer
        rows = np.full((n_shards, n_f, k), -1, dtype=np.int64)
        for s in range(n_shards):
            for f in range(n_f):
                picked = np.sort(rng.choice(1000, size=counts[s, f], replace=False))[::-1] + 1000 * s
                rows[s, f, :counts[s, f]] = picked
        want = np.full((n_f, k), -1, dtype=np.int64)
        want_count = np.zeros(n_f, dtype=np.int32)
        for f in range(n_f):
            merged = [int(r) for s in range(n_shards - 1, -1, -1) for r in rows[s, f, :counts[s, f]]][:k]
            want[f, :len(merged)] = merged
            want_count[f] = len(merged)
        got, got_count = sharded.combine_rescans(torch.as_tensor(rows), torch.as_tensor(counts), k)
        assert np.array_equal(got.numpy(), want) and np.array_equal(got_count.numpy(), want_count)
        assert got_count.dtype == torch.int32 and got.dtype == torch.int64
