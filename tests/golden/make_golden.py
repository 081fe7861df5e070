"""Mints the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

Every expected value below is produced by the reference's own functions (numba 0.65.0, numpy 2.3,
CPython 3.12 - "as executed here").  Hash-order dependent quantities (column ids,
per-row n-gram order: match_maker.py:144-147, :172-174) are stored explicitly so that the fixtures
reproduce the reference bit-for-bit in any process.

Outputs
  example_titles.npz    transformed truth/test titles + title ids of the example data set
  matchmaker_example.npz  encoded index (vocab, CSR in set order), reference sums / mx / jaccard rows,
                          reference candidate lists for the first N_QUERIES test rows at k=10 and k=100
  topk_vectors.npz      fast_arg_top_k known answers
  pairs.npz             fast_levenshtein_ratio + construct_features inputs and reference outputs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

N_QUERIES = 1000
ALPHABET = '- abcdefghijklmnopqrstuvwxyz0123456789'
ENC = {ch: i for i, ch in enumerate(ALPHABET)}


def encode(title):
    out = np.zeros(255, dtype=np.uint8)
    codes = [ENC[ch] for ch in title[:255]]
    out[:len(codes)] = codes
    return out


def main():
    ref = ref_import.import_reference()
    c = ref.constants
    rng = np.random.default_rng(20240601)

    truth = ref.common.get_ground_truth()
    test = ref.common.get_test_data()
    truth_titles = list(truth[c.COLUMN_TRANSFORMED_TITLE])
    test_titles = list(test[c.COLUMN_TRANSFORMED_TITLE])
    np.savez_compressed(
        os.path.join(HERE, 'example_titles.npz'),
        truth_titles=np.array(truth_titles), truth_title_ids=truth[c.COLUMN_TITLE_ID].to_numpy(np.int64),
        test_titles=np.array(test_titles), test_index=test[c.COLUMN_TEST_INDEX].to_numpy(np.int64))

    # ---------------- MatchMaker (match_maker.py) ----------------
    data = test.iloc[:N_QUERIES].copy()
    mm = ref.match_maker.MatchMaker(data.copy(), truth.copy(), 100)
    vocab = [mm.n_grams_decoding[i] for i in range(len(mm.n_grams_decoding))]
    encoding = mm.n_grams_encoding

    def csr(rows):
        ptr = np.zeros(len(rows) + 1, dtype=np.int64)
        cols = []
        for i, value in enumerate(rows):
            cols.extend(encoding[x] for x in value)       # the set's own iteration order (:172-174)
            ptr[i + 1] = len(cols)
        return ptr, np.array(cols, dtype=np.uint16)

    t_ptr, t_cols = csr(list(truth[c.COLUMN_N_GRAMS]))
    q_ptr, q_cols = csr(list(data[c.COLUMN_N_GRAMS]))
    w64 = np.array([mm._get_idf_given_index(i) for i in range(len(vocab))], dtype=np.float64)
    mx = np.array([sum([mm._get_idf_given_index(r) for r in mm.matrix_non_zero_columns[q]])
                   for q in range(N_QUERIES)], dtype=np.float64)                      # :197
    jac_queries = np.array([0, 1, 2, 7], dtype=np.int64)
    jac = np.stack([ref.match_maker.fast_jaccard(
        mm.number_of_truth_titles, mx[q], mm.matrix_non_zero_columns[q],
        mm.matrix_truth_non_zero_columns_and_values, mm.sums_matrix_truth) for q in jac_queries])
    title_ids = truth[c.COLUMN_TITLE_ID].to_numpy(np.int64)
    id_to_row = {int(t): i for i, t in enumerate(title_ids)}
    top = {}
    for k in (10, 100):
        mm.top_n = k
        top[k] = np.array([[id_to_row[t] for t in mm.get_closest_matches(q)] for q in range(N_QUERIES)],
                          dtype=np.int32)
    np.savez_compressed(
        os.path.join(HERE, 'matchmaker_example.npz'),
        vocab=np.array(vocab), w64=w64, t_ptr=t_ptr, t_cols=t_cols, q_ptr=q_ptr, q_cols=q_cols,
        sums=mm.sums_matrix_truth.astype(np.float32), mx=mx, jac_queries=jac_queries, jac=jac,
        top10_rows=top[10], top100_rows=top[100], title_ids=title_ids)

    # ---------------- fast_arg_top_k known answers ----------------
    fatk = ref.match_maker.fast_arg_top_k
    cases = [
        (np.array([0.9, 0.5, 0.5, 0.5]), 2), (np.zeros(8), 3), (np.array([0, 0.3, 0, 0.2, 0, 0]), 3),
        (np.array([0.5, 0.5000004, 0.4999996, 0.1, 0.9]), 2), (np.array([0.1, 0.2, 0.3, 0.4, 0.5]), 2),
        (np.array([0.1, 0.2]), 3),
    ]
    for n, k in ((50, 5), (200, 10), (1000, 100), (300, 100), (64, 10)):
        for mode in range(4):
            v = rng.random(n)
            if mode == 1:
                v = np.round(v, 1)                      # massive ties
            if mode == 2:
                v[rng.random(n) < 0.9] = 0.0            # fewer than k positives
            if mode == 3:
                v = 0.5 + (rng.integers(-3, 4, n) * 4e-7)   # near-ties inside the 1e-6 band
            cases.append((v.astype(np.float64), k))
    tk = {}
    for i, (v, k) in enumerate(cases):
        tk[f'v{i}'] = v.astype(np.float64)
        tk[f'k{i}'] = np.int64(k)
        tk[f'r{i}'] = fatk(v.astype(np.float64), k).astype(np.int64)
    tk['n_cases'] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, 'topk_vectors.npz'), **tk)

    # ---------------- pair scoring (feature_engineering.py) ----------------
    fe = ref.feature_engineering
    flr = fe.fast_levenshtein_ratio

    def arr(x):
        return np.array(x, dtype=np.uint8)

    ratio_pairs = [
        (arr([2] * 29 + [3] * 21), arr([2] * 29 + [4] * 21)), (arr([2] * 128), arr([3] * 128)),
        (arr([2] * 127), arr([3] * 128)), (arr([2] * 130), arr([3] * 130)), (arr([2] * 200), arr([3] * 100)),
        (arr([2] * 255), arr([3] * 255)), (arr([2] * 255), arr([3] * 1)), (arr([2] * 200), arr([3] * 200)),
        (arr([]), arr([1, 2])), (encode('coolblue bv')[:11], encode('coolblue bv')[:11]),
    ]
    for it in range(4000):
        big = it % 4 == 0
        la = int(rng.integers(1, 256 if big else 48))
        lb = int(rng.integers(1, 256 if big else 48))
        alpha = int(rng.integers(2, 12))
        a = rng.integers(1, 1 + alpha, la).astype(np.uint8)
        b = rng.integers(1, 1 + alpha, lb).astype(np.uint8)
        if it % 3 == 0:
            m = min(la, lb)
            b[:m] = a[:m]
            if m > 2:
                b[int(rng.integers(0, m))] = 37
        ratio_pairs.append((a, b))
    r_la = np.array([len(a) for a, b in ratio_pairs], dtype=np.int32)
    r_lb = np.array([len(b) for a, b in ratio_pairs], dtype=np.int32)
    r_a = np.zeros((len(ratio_pairs), 255), dtype=np.uint8)
    r_b = np.zeros((len(ratio_pairs), 255), dtype=np.uint8)
    r_out = np.zeros(len(ratio_pairs), dtype=np.uint8)
    for i, (a, b) in enumerate(ratio_pairs):
        r_a[i, :len(a)] = a
        r_b[i, :len(b)] = b
        r_out[i] = flr(a, b)

    wc = ref.common.get_words_counter(truth)
    pairs, counts = [], []
    # candidate-like pairs: a test title against its reference top-100 candidates (sampled)
    for q in range(0, N_QUERIES, 5):
        for t in rng.choice(top[100][q], 8, replace=False):
            pairs.append((test_titles[q], truth_titles[int(t)]))
    pairs += [('coolblu bv', 'coolblue bv'), ('abc', 'xyz'), ('feld s ullivan limited', 'field sullivan limited')]
    for a, b in pairs:
        cnt = np.zeros(15, dtype=np.uint32)
        ws = [wc.get(w, 0) for w in b.split()][:15]
        cnt[:len(ws)] = ws
        counts.append(cnt)
    words = ['ab', 'cde', 'fghi', 'jk', 'lmnop', 'q', 'rst', 'uv', 'wxyz', 'a1', 'b22', 'c333', 'dd', 'ee',
             'ffg', 'hh', 'ii', 'jj', 'kk', 'll', 'limited', 'ltd', 'bv', 'holdings', 'international']
    for nw in range(1, 21):
        for rep in range(12):
            t = ' '.join(rng.choice(words, nw))
            q = ' '.join(rng.choice(words, max(1, nw - 1 + int(rng.integers(-1, 2)))))
            if len(t) > 255 or len(q) > 255:
                continue
            pairs.append((q, t))
            cnt = rng.integers(1, 3000, 15).astype(np.uint32)
            if rep % 4 == 0:
                cnt[int(rng.integers(0, 15))] = 0          # df 0 -> +inf idf
            if rep % 6 == 5:
                cnt[:] = 0
            cnt[min(nw, 15):] = 0
            counts.append(cnt)
    for it in range(60):                                   # long titles: uint8 wrap region, long words
        la, lb = int(rng.integers(100, 256)), int(rng.integers(100, 256))
        chars = list('abc d') if it % 2 == 0 else list('ab')

        def mk(n):
            s = ' '.join(''.join(rng.choice(chars, n)).split())
            return s if len(s) >= 3 else 'abc'
        pairs.append((mk(la), mk(lb)))
        counts.append(rng.integers(1, 3000, 15).astype(np.uint32))
    p_la = np.array([len(a) for a, b in pairs], dtype=np.uint8)
    p_lb = np.array([len(b) for a, b in pairs], dtype=np.uint8)
    p_a = np.vstack([encode(a) for a, b in pairs])
    p_b = np.vstack([encode(b) for a, b in pairs])
    p_counts = np.vstack(counts).astype(np.uint32)
    feats = np.zeros((len(pairs), 66), dtype=np.float32)
    with np.errstate(all='ignore'):
        fe.construct_features(p_la, p_lb, p_a, p_b, p_counts, np.uint8(1), np.uint32(30000),
                              np.zeros(66, dtype=np.uint8), feats)
    np.savez_compressed(
        os.path.join(HERE, 'pairs.npz'),
        ratio_a=r_a, ratio_b=r_b, ratio_la=r_la, ratio_lb=r_lb, ratio_out=r_out,
        feat_la=p_la, feat_lb=p_lb, feat_a=p_a, feat_b=p_b, feat_counts=p_counts, feat_n_truth=np.uint32(30000),
        feat_out=feats, feat_titles=np.array([a for a, b in pairs]), feat_truths=np.array([b for a, b in pairs]))
    for name in sorted(os.listdir(HERE)):
        if name.endswith('.npz'):
            print(name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == '__main__':
    main()
