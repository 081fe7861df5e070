"""GPU: BASELINE.json's full sizes.  C3 (100k x 500k, top-10) is checked through size-independent properties
(strictly descending rows, idempotence, host-buffer path == device-buffer path, every returned row reaches
the reference threshold) plus an oracle comparison of sampled queries; C4 through sampled pairs of a 10M-pair
batch (the full 100M run is bench_pairs.py)."""
import numpy as np
import pytest

from tests.conftest import features_equal, oracle_index_from_encoded

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c3():
    from doppelspeller_b200 import encode, synthetic
    truth = synthetic.generate_truth_titles(500_000)
    test, _ = synthetic.generate_test_titles(truth, 100_000)
    return encode.encode_canonical(test, truth)


def test_c3_properties_and_sampled_parity(c3):
    import torch
    from doppelspeller_b200.index import TruthIndex
    from oracle import oracle
    k, n_truth, n_q = 10, 500_000, 100_000
    index = TruthIndex(c3['t_ptr'], c3['t_cols'], c3['idf64'])
    rows, count, kth, flags = index.topn(c3['q_ptr'], c3['q_cols'], k, with_details=True)
    assert rows.shape == (n_q, k) and (count == k).all()
    assert ((rows >= 0) & (rows < n_truth)).all()
    assert (np.diff(rows, axis=1) < 0).all()                       # descending truth row index (match_maker.py:71)
    # idempotence and device-resident inputs
    q_ptr, q_cols = torch.as_tensor(c3['q_ptr']).cuda(), torch.as_tensor(c3['q_cols']).cuda()
    rows_dev, count_dev = index.topn(q_ptr, q_cols, k)
    assert np.array_equal(rows_dev.cpu().numpy(), rows)
    # sampled queries against the oracle, plus the threshold property on their exact scores
    sample = np.sort(np.random.default_rng(1).choice(n_q, 400, replace=False))
    oracle_index = oracle_index_from_encoded(c3)
    want_rows, want_count, want_kth = oracle.topn(oracle_index, k, queries=sample)
    assert np.array_equal(rows[sample], want_rows)
    assert np.array_equal(kth[sample], want_kth)
    for q in sample[:40]:
        scores = oracle.fast_jaccard(oracle_index, int(q))
        thr = float(kth[q]) - float(np.float32(1e-6))
        assert (scores[rows[q]] >= thr).all()
        qualifying = np.nonzero(scores >= thr)[0]
        assert np.array_equal(rows[q], qualifying[::-1][:k])       # the k highest qualifying rows
    assert (flags[(flags & 2) != 0] & 2).all()


def test_c4_sampled_pairs():
    import torch
    from collections import Counter
    from doppelspeller_b200 import feature_engineering as fe
    from doppelspeller_b200 import synthetic
    from oracle import oracle
    rng = np.random.default_rng(synthetic.PAIRS_SEED)
    n_titles, n_pairs = 100_000, 10_000_000
    truth = synthetic.generate_truth_titles(n_titles - n_titles // 10, seed=31) + synthetic.generate_long_titles(n_titles // 10, seed=32)
    test, source = synthetic.generate_test_titles(truth, n_titles, seed=33, matched=1.0)
    idx_a = rng.integers(0, n_titles, n_pairs).astype(np.int32)
    idx_b = np.where(rng.random(n_pairs) < 0.5, source[idx_a], rng.integers(0, n_titles, n_pairs)).astype(np.int32)
    counter = Counter(w for t in truth for w in set(t.split()))
    counts = np.zeros((n_titles, 15), dtype=np.uint32)
    for i, t in enumerate(truth):
        ws = [counter[w] for w in t.split()[:15]]
        counts[i, :len(ws)] = ws
    table_a, table_b = fe.encode_titles(test), fe.encode_titles(truth)
    dev = lambda x: torch.as_tensor(x).cuda()   # noqa: E731
    feats = fe.construct_features_pairs((dev(table_a[0]), dev(table_a[1])), (dev(table_b[0]), dev(table_b[1])),
                                        dev(counts.view(np.int32)), dev(idx_a), dev(idx_b), fe.SPACE_CODE, n_titles)
    assert feats.shape == (n_pairs, 66)
    sample = np.sort(rng.choice(n_pairs, 20_000, replace=False))
    got = feats[torch.as_tensor(sample).cuda()].cpu().numpy()
    la = np.array([len(test[i]) for i in idx_a[sample]], dtype=np.uint8)
    lb = np.array([len(truth[i]) for i in idx_b[sample]], dtype=np.uint8)
    pa = np.vstack([fe.encode_title(test[i]) for i in idx_a[sample]])
    pb = np.vstack([fe.encode_title(truth[i]) for i in idx_b[sample]])
    want = oracle.construct_features(la, lb, pa, pb, counts[idx_b[sample]], fe.SPACE_CODE, n_titles)
    assert features_equal(got, want)
    # feature 4 is fast_levenshtein_ratio(title, truth): the standalone batched kernel must agree with it
    ratio = fe.fast_levenshtein_ratio_batch(pa, pb, la, lb)
    assert np.array_equal(ratio.astype(np.float32), got[:, 4])
    # size-independent sanity over the whole batch: lengths echo the inputs, ratios are percentages
    head = feats[:, :6].cpu().numpy()
    assert np.array_equal(head[:, 0], np.diff(table_a[1])[idx_a].astype(np.float32))
    assert np.array_equal(head[:, 1], np.diff(table_b[1])[idx_b].astype(np.float32))
    assert ((head[:, 4] >= 0) & (head[:, 4] <= 100) & (head[:, 5] >= 0) & (head[:, 5] <= 100)).all()
