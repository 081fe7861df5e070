"""The synthetic workload of bench.py is pinned to the example data's statistics: the numbers below are
what `config` reports in every bench line; the kernels' work (postings touched per query) scales with them."""
import os

import numpy as np

from tests.conftest import GOLDEN


def _example():
    raw = np.load(os.path.join(GOLDEN, 'example_titles.npz'))
    return [str(t) for t in raw['truth_titles']], [str(t) for t in raw['test_titles']]


def test_generator_matches_the_example_density():
    from doppelspeller_b200 import encode, synthetic
    truth, test = _example()
    want = synthetic.workload_statistics(encode.encode_canonical(test, truth))
    assert abs(want['top_trigram_df_share'] - 0.2745) < 1e-3 and abs(want['postings_hit_per_query_over_n'] - 0.890) < 2e-3
    titles = synthetic.generate_truth_titles(30000)
    queries, source = synthetic.generate_test_titles(titles, 10000)
    got = synthetic.workload_statistics(encode.encode_canonical(queries, titles))
    assert abs(got['top_trigram_df_share'] - want['top_trigram_df_share']) < 0.02            # top trigram in 27 % +- 2 of the titles
    assert abs(got['postings_hit_per_query_over_n'] - want['postings_hit_per_query_over_n']) < 0.05   # ~0.9 N postings per query
    assert abs(got['mean_trigrams_per_truth_title'] - want['mean_trigrams_per_truth_title']) < 1.0
    assert 3 <= got['median_trigram_df'] <= 8                                               # example: 4
    lengths = np.array([len(t) for t in titles])
    assert abs(lengths.mean() - 23.4) < 1.0 and np.percentile(lengths, 95) < 50 and lengths.min() >= 3
    assert len(set(titles)) > 0.995 * len(titles)                                           # the example's titles are 99.9 % distinct
    assert 0.55 < (source >= 0).mean() < 0.65                                               # 60 % of the test titles are edited truth titles
    # deterministic
    assert synthetic.generate_truth_titles(500) == synthetic.generate_truth_titles(500)


def test_tiled_example_titles_keep_their_density():
    from doppelspeller_b200 import encode, synthetic
    truth, test = _example()
    tiled = synthetic.tile_titles(truth, 90000, synthetic.TRUTH_SEED)
    queries = synthetic.tile_titles(test, 15000, synthetic.TEST_SEED)
    assert tiled[:len(truth)] == truth
    stats = synthetic.workload_statistics(encode.encode_canonical(queries, tiled))
    assert abs(stats['top_trigram_df_share'] - 0.2745) < 0.01
    assert 0.75 < stats['postings_hit_per_query_over_n'] < 0.92
    assert len(set(tiled)) > 0.9 * len(tiled)


def test_restricted_oracle_index_gives_the_same_candidates():
    """bench.py checks C5 (10M truth rows) against an oracle index that only holds the posting lists its query sample
    touches: same candidate lists as the full index."""
    import bench
    from doppelspeller_b200 import encode, synthetic
    from oracle import oracle
    truth = synthetic.generate_truth_titles(20000, seed=3)
    test, _ = synthetic.generate_test_titles(truth, 2000, seed=4)
    enc = encode.encode_canonical(test, truth)
    sample = bench.cpu_sample(2000, 150)
    full, part = bench.oracle_index(enc), bench.oracle_index(enc, queries=sample)
    assert part['post_rows'].shape[0] < full['post_rows'].shape[0]
    for k in (10, 100):
        want = oracle.topn(full, k, queries=sample)
        got = oracle.topn(part, k, queries=sample)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
