"""TEST INFRASTRUCTURE ONLY: time-boxed shape fuzzer of the EMULATED kernels (tests/emu/cuda_emu.h), meant to run under
the sanitizers.  Every case is checked against the CPU oracle; any sanitizer report aborts the process.

    python tests/emu/build.py --asan
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    DOPPELSPELLER_B200_LIB=$PWD/tests/emu/_build/libds_emu_asan.so python tests/emu/fuzz.py --minutes 30 --seed 1

Shapes aim at the boundaries the kernels index by: 4,096-row posting blocks and the 8,192-row switch to the posting form,
128-row groups, 32-column query groups, 64-posting pieces, candidate-buffer capacities (256 / 1,024 / 4,096), retained
lists, 32 / 64 / 255-character strings, the uint8 wrap region, 15 / 16 / 17-word titles, runs of 1..40 pairs per title.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

if 'libds_emu' not in os.path.basename(os.environ.get('DOPPELSPELLER_B200_LIB', '')):
    raise SystemExit('fuzz.py drives the emulated library only: set DOPPELSPELLER_B200_LIB (see the header)')

import torch  # noqa: E402

torch.cuda.is_available = lambda: True
torch.cuda.current_device = lambda: 0
torch.cuda.current_stream = lambda device=None: type('S', (), {'cuda_stream': 0})()

from doppelspeller_b200 import _native as nat  # noqa: E402
from doppelspeller_b200 import common, encode, predict, sharded  # noqa: E402
from doppelspeller_b200 import feature_engineering as fe  # noqa: E402
from doppelspeller_b200.index import TruthIndex, topn_merge  # noqa: E402
from oracle import oracle  # noqa: E402

ALPHABET = np.frombuffer(b' abcdefghijklmnopqrstuvwxyz0123456789', dtype=np.uint8)


def oracle_index(enc):
    return oracle.finish_index(dict(n_truth=int(enc['t_ptr'].shape[0]) - 1, w64=enc['idf64'], w32=enc['idf64'].astype(np.float32),
                                    t_ptr=enc['t_ptr'], t_cols=enc['t_cols'].astype(np.int32), q_ptr=enc['q_ptr'],
                                    q_cols=enc['q_cols'].astype(np.int32)))


def csr(sets, dtype=np.uint16):
    ptr = np.zeros(len(sets) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in sets], out=ptr[1:])
    cols = np.fromiter((c for s in sets for c in s), dtype=dtype, count=int(ptr[-1]))
    return ptr, cols


def case_topn(rng):
    n_truth = int(rng.choice([1, 2, 31, 127, 129, 255, 4095, 4097, 8191, 8192, 8193, 8320, 12287, 12289, 16385, 20480]))
    n_truth = max(1, n_truth + int(rng.integers(-2, 3)))
    n_vocab = int(rng.choice([2, 3, 12, 33, 60, 500, 5000, 40000, 65535]))
    n_q = int(rng.choice([1, 3, 31, 33, 64, 150]))
    k = int(rng.choice([1, 2, 5, 10, 32, 37, 64, 100, 250, 512, 1024]))
    skew = float(rng.choice([0.0, 0.7, 1.1, 1.6]))
    weights = 1.0 / np.arange(1, n_vocab + 1) ** skew
    weights /= weights.sum()

    def rows_of(count, max_len):
        lengths = rng.integers(0, min(max_len, n_vocab) + 1, count)
        drawn = rng.choice(n_vocab, size=int(lengths.sum()), p=weights)
        return [sorted(set(part.tolist())) for part in np.split(drawn, np.cumsum(lengths)[:-1])]
    truth = rows_of(n_truth, int(rng.choice([4, 30, 120])))
    mode = int(rng.integers(0, 5))
    if mode == 0 and n_truth > 8:                        # heavy duplication: ties, overflowing candidate buffers
        copies = rng.integers(0, n_truth, max(1, int(n_truth * rng.choice([0.05, 0.3, 0.8]))))
        source = int(rng.integers(0, n_truth))
        for r in copies:
            truth[int(r)] = truth[source]
    if mode == 1 and n_vocab > 2:                        # a column in EVERY row: idf 0 (weightless), others rare
        truth = [sorted(set(s) | {0}) for s in truth]
    if not any(truth):
        truth[0] = [0]
    queries = rows_of(n_q, int(rng.choice([5, 40, 250])))
    queries[0] = truth[int(rng.integers(0, n_truth))]
    if n_q > 2:
        queries[1] = []
    import math
    df = np.zeros(n_vocab)
    for s in truth:
        for c in s:
            df[c] += 1
    idf = np.array([math.log(n_truth / d) if d > 0 else 0.0 for d in df])
    idf[df == 0] = idf.max()
    t_ptr, t_cols = csr(truth)
    q_ptr, q_cols = csr(queries)
    enc = dict(idf64=idf, t_ptr=t_ptr, t_cols=t_cols, q_ptr=q_ptr, q_cols=q_cols)
    want_rows, want_count, want_kth = oracle.topn(oracle_index(enc), k)
    index = TruthIndex(t_ptr, t_cols, idf)
    rows, count, kth, flags = index.topn(q_ptr, q_cols, k, with_details=True)
    index.close()
    assert np.array_equal(count, want_count), 'count'
    assert np.array_equal(rows, want_rows), 'rows'
    assert np.array_equal(kth, want_kth), 'kth'
    # the same through truth shards (local -> merge -> rescan), with thresholds shared through peer arrays half the time
    n_shards = int(rng.integers(2, 6))
    if n_truth >= n_shards:
        offs = sharded.shard_offsets(n_truth, n_shards)
        shards = []
        for r in range(n_shards):
            ptr, cols = sharded.slice_truth_csr(t_ptr, t_cols, int(offs[r]), int(offs[r + 1]))
            shards.append(TruthIndex(ptr, cols, idf, row_offset=int(offs[r]), n_total=n_truth))
        share = bool(rng.integers(0, 2))
        thetas = None
        if share:
            lib = nat.lib
            lib.ds_emu_malloc.restype = __import__('ctypes').c_void_p
            import ctypes
            thetas = []
            for _ in shards:
                address = lib.ds_emu_malloc(ctypes.c_size_t(max(8, 8 * n_q)))
                array = np.frombuffer((ctypes.c_char * (8 * n_q)).from_address(address), dtype=np.float64, count=n_q)
                array[...] = 0.0
                thetas.append((address, array))
        local = {}
        for r in rng.permutation(n_shards):
            r = int(r)
            if share:
                peers = [thetas[p][0] for p in range(n_shards) if p != r]
                local[r] = shards[r].topn_local(q_ptr, q_cols, k, theta_own=thetas[r][1], theta_peers=peers)
            else:
                local[r] = shards[r].topn_local(q_ptr, q_cols, k)
        all_score = np.stack([local[r][0] for r in range(n_shards)])
        all_row = np.stack([local[r][1] for r in range(n_shards)])
        m_rows, m_count, m_kth, thr, m_flags = topn_merge(all_score, all_row, k, n_truth, q_mx=local[0][2])
        flagged = np.nonzero(m_flags & 1)[0]
        if flagged.size:
            per_rows, per_count = [], []
            for s in shards:
                lr = np.full((n_q, k), -1, dtype=np.int64)
                lc = np.zeros(n_q, dtype=np.int32)
                s.topn_rescan(q_ptr, q_cols, local[0][2], thr, m_flags, k, lr, lc)
                per_rows.append(lr[flagged])
                per_count.append(lc[flagged])
            fixed, fixed_count = sharded.combine_rescans(torch.as_tensor(np.stack(per_rows)), torch.as_tensor(np.stack(per_count)), k)
            m_rows[flagged] = fixed.numpy()
            m_count[flagged] = fixed_count.numpy()
        for s in shards:
            s.close()
        if share:
            import ctypes
            for address, _ in thetas:
                nat.lib.ds_emu_free(ctypes.c_void_p(address))
        assert np.array_equal(m_count, want_count), 'sharded count'
        assert np.array_equal(m_rows, want_rows), 'sharded rows'
    return f'topn n={n_truth} V={n_vocab} q={n_q} k={k} mode={mode}'


def random_codes(rng, n, alphabet_size, wild=0.0):
    a = (rng.integers(0, 1 << 30, (n, 255)) % alphabet_size).astype(np.uint8)
    if wild > 0:
        a[rng.random(n) < wild, int(rng.integers(0, 8))] = 200
    return a


def case_pairs_padded(rng):
    n = int(rng.choice([1, 31, 33, 257, 1500]))
    lengths = [lambda m: rng.integers(0, 256, m), lambda m: rng.integers(1, 33, m), lambda m: rng.integers(30, 70, m),
               lambda m: rng.integers(120, 136, m), lambda m: np.full(m, int(rng.choice([0, 1, 32, 33, 64, 65, 127, 128, 255])))]
    la = lengths[int(rng.integers(0, len(lengths)))](n).astype(np.uint8)
    lb = lengths[int(rng.integers(0, len(lengths)))](n).astype(np.uint8)
    alpha = int(rng.choice([2, 3, 8, 38]))
    a = random_codes(rng, n, alpha, wild=0.02 if rng.random() < 0.3 else 0.0)
    b = random_codes(rng, n, alpha)
    similar = rng.random(n) < 0.6
    b[similar] = a[similar]
    for j in range(int(rng.integers(0, 6))):
        b[np.arange(n), rng.integers(0, 255, n)] = rng.integers(0, alpha, n)
    got, dist = fe.fast_levenshtein_ratio_batch(a, b, la, lb, with_distance=True)
    want = oracle.indel_ratio_u8_batch(a, b, la, lb)
    assert np.array_equal(got, want), 'indel ratio'
    counts = rng.integers(0, 5000, size=(n, 15)).astype(np.uint32)
    counts[rng.random((n, 15)) < 0.1] = 0
    # word structure: sprinkle spaces (code 1) so that titles hold 1..20 words
    spaces = rng.random((n, 255)) < float(rng.choice([0.02, 0.12, 0.3]))
    a2, b2 = a.copy(), b.copy()
    a2[spaces] = 1
    b2[spaces & (rng.random((n, 255)) < 0.9)] = 1
    a2, b2 = np.minimum(a2, 37), np.minimum(b2, 37)
    n_truth = int(rng.choice([1, 3, 30000, 4000000]))
    got_f = fe.construct_features(la, lb, a2, b2, counts, fe.SPACE_CODE, n_truth)
    want_f = oracle.construct_features(la, lb, a2, b2, counts, fe.SPACE_CODE, n_truth)
    same_nan = np.isnan(got_f) == np.isnan(want_f)
    exact = (got_f[:, :36] == want_f[:, :36]) | (np.isnan(got_f[:, :36]) & np.isnan(want_f[:, :36]))
    with np.errstate(all='ignore'):
        close = np.isclose(got_f[:, 36:], want_f[:, 36:], rtol=1e-6, atol=0, equal_nan=True)
    assert same_nan.all() and exact.all() and close.all(), 'features'
    return f'pairs n={n} alpha={alpha}'


def random_titles(rng, n, low, high, words=True):
    out = []
    for _ in range(n):
        length = int(rng.integers(low, high + 1))
        codes = ALPHABET[rng.integers(1, len(ALPHABET), length)].copy()
        if words and length > 4:
            codes[rng.random(length) < float(rng.choice([0.05, 0.15, 0.35]))] = 32
        text = ' '.join(codes.tobytes().decode('ascii').split())
        out.append(text if len(text) >= 3 else (text + '000')[:3])
    return out


def case_tables(rng):
    """table forms: candidate runs (k_indel_groups), the sorted class pipeline, k_feature_words, the pre-match cascade"""
    n_a = int(rng.choice([1, 7, 60]))
    n_b = int(rng.choice([5, 90, 700]))
    long_share = float(rng.choice([0.0, 0.1, 0.5]))
    titles_a = random_titles(rng, n_a, 3, 60) if rng.random() > long_share else random_titles(rng, n_a, 60, 255)
    titles_b = random_titles(rng, n_b, 3, 70)
    for i in range(0, n_b, 3):                           # near copies: the high-ratio region and the token-sort branch
        src = titles_a[int(rng.integers(0, n_a))]
        titles_b[i] = src[:-1] if i % 2 else ' '.join(reversed(src.split()))
        if len(titles_b[i]) < 3:
            titles_b[i] = 'abc'
    run = int(rng.choice([1, 3, 4, 10, 33, 40]))
    idx_a = np.repeat(np.arange(n_a, dtype=np.int32), run)
    if rng.random() < 0.3:
        rng.shuffle(idx_a)                               # not runs at all
    idx_b = rng.integers(0, n_b, len(idx_a)).astype(np.int32)
    ca, cb = fe.encode_titles(titles_a), fe.encode_titles(titles_b)
    got = np.empty(len(idx_a), np.uint8)
    nat.check(nat.lib.ds_indel_ratio_pairs(nat.ptr(ca[0]), nat.ptr(ca[1]), n_a, nat.ptr(cb[0]), nat.ptr(cb[1]), n_b, nat.ptr(idx_a),
                                           nat.ptr(idx_b), len(idx_a), nat.ptr(got), None, None))
    a = np.vstack([fe.encode_title(titles_a[i]) for i in idx_a])
    b = np.vstack([fe.encode_title(titles_b[i]) for i in idx_b])
    la = np.array([len(titles_a[i]) for i in idx_a], np.uint8)
    lb = np.array([len(titles_b[i]) for i in idx_b], np.uint8)
    assert np.array_equal(got, oracle.indel_ratio_u8_batch(a, b, la, lb)), 'table indel'
    from doppelspeller_b200.pipeline import truth_word_counts
    counts = truth_word_counts(titles_b)
    feats = fe.construct_features_pairs(ca, cb, counts, idx_a, idx_b, fe.SPACE_CODE, n_b)
    want = oracle.construct_features(la, lb, a, b, counts[idx_b], fe.SPACE_CODE, n_b)
    exact = (feats[:, :36] == want[:, :36]) | (np.isnan(feats[:, :36]) & np.isnan(want[:, :36]))
    with np.errstate(all='ignore'):
        close = np.isclose(feats[:, 36:], want[:, 36:], rtol=1e-6, atol=0, equal_nan=True)
    assert exact.all() and close.all() and (np.isnan(feats) == np.isnan(want)).all(), 'table features'
    xs, ys = [titles_a[i] for i in idx_a], [titles_b[i] for i in idx_b]
    assert np.array_equal(common.levenshtein_ratio_batch(xs, ys), np.array([oracle.levenshtein_ratio(x, y) for x, y in zip(xs, ys)])), 'ratio'
    assert np.array_equal(predict.get_levenshtein_ratios(xs, ys), np.array([oracle.prematch_ratio(x, y) for x, y in zip(xs, ys)])), 'prematch'
    return f'tables a={n_a} b={n_b} run={run}'


def case_encoder(rng):
    n_truth = int(rng.choice([1, 33, 700, 5000]))
    n_q = int(rng.choice([0, 1, 40, 300]))
    high = int(rng.choice([3, 12, 60, 255]))
    truth = random_titles(rng, n_truth, 3, high, words=bool(rng.integers(0, 2)))
    test = random_titles(rng, n_q, 3, high)
    want = encode.encode_canonical(test, truth)
    t_bytes, t_off = encode.title_table(truth)
    q_bytes, q_off = encode.title_table(test)
    import ctypes
    max_vocab = int(nat.lib.ds_encode_max_vocab())
    t_ptr, q_ptr = np.empty(n_truth + 1, np.int64), np.empty(n_q + 1, np.int64)
    t_cols, q_cols = np.empty(max(1, int(t_off[-1])), np.uint16), np.empty(max(1, int(q_off[-1])), np.uint16)
    idf64, vocab = np.empty(max_vocab, np.float64), np.empty(max_vocab, np.int32)
    n_vocab, t_nnz, q_nnz = ctypes.c_int32(0), ctypes.c_int64(0), ctypes.c_int64(0)
    nat.check(nat.lib.ds_encode_trigrams(nat.ptr(t_bytes), nat.ptr(t_off), n_truth, nat.ptr(q_bytes), nat.ptr(q_off), n_q, nat.ptr(t_ptr),
                                         nat.ptr(t_cols), nat.ptr(q_ptr), nat.ptr(q_cols), nat.ptr(idf64), nat.ptr(vocab),
                                         ctypes.byref(n_vocab), ctypes.byref(t_nnz), ctypes.byref(q_nnz), 0, None))
    assert np.array_equal(t_ptr, want['t_ptr']) and np.array_equal(t_cols[:t_nnz.value], want['t_cols']), 'truth csr'
    if n_q:
        assert np.array_equal(q_ptr, want['q_ptr']) and np.array_equal(q_cols[:q_nnz.value], want['q_cols']), 'query csr'
    assert np.array_equal(vocab[:n_vocab.value], want['vocab_codes']), 'vocab'
    assert np.array_equal(idf64[:n_vocab.value].view(np.uint64), want['idf64'].view(np.uint64)), 'idf'
    # title features of the same table
    codes = np.empty(max(1, t_bytes.shape[0]), np.uint8)
    counts = np.empty((n_truth, 15), np.uint32)
    nat.check(nat.lib.ds_title_features(nat.ptr(t_bytes), nat.ptr(t_off), n_truth, nat.ptr(codes), nat.ptr(counts), 0, None))
    from doppelspeller_b200.pipeline import truth_word_counts
    want_codes, _ = fe.encode_titles(truth)
    assert np.array_equal(codes[:want_codes.shape[0]], want_codes), 'codes'
    assert np.array_equal(counts, truth_word_counts(truth)), 'word counts'
    # transform_title of raw strings built around the same titles
    raw = [t.upper().replace(' ', '  -') if i % 2 else '\t' + t + 'é中!' for i, t in enumerate(truth[:200])]
    raw += ['', '-', 'x' * 300, 'Ångström  & co.', 'a\x1cb']
    assert common.transform_titles(raw) == [oracle.transform_title(t) for t in raw], 'transform'
    return f'encoder n={n_truth} q={n_q} len<={high}'


def case_gbdt(rng):
    from doppelspeller_b200 import gbdt
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from tests.test_gpu_parity import _random_forest
    n_trees = int(rng.choice([0, 1, 3, 4, 5, 70, 400]))
    depth = int(rng.choice([1, 3, 5, 9, 13]))
    n_features = int(rng.choice([1, 7, 66, 200]))
    model = gbdt.GbdtModel.from_trees(_random_forest(rng, n_trees, n_features, depth), base_margin=float(rng.normal()),
                                      transform=gbdt.LOGISTIC)
    n = int(rng.choice([1, 255, 257, 3000]))
    x = rng.normal(0, 50, size=(n, n_features)).astype(np.float32)
    x[rng.random(x.shape) < 0.08] = np.nan
    x[rng.random(x.shape) < 0.01] = np.inf
    got = model.predict(x)
    want = oracle.gbdt_predict(x, model.nodes, model.tree_offsets, model.base_margin, logistic=True)
    assert np.allclose(got, want, rtol=1e-6, atol=0), 'gbdt'
    return f'gbdt trees={n_trees} depth={depth} features={n_features} rows={n}'


CASES = [(case_topn, 5), (case_pairs_padded, 2), (case_tables, 2), (case_encoder, 1), (case_gbdt, 1)]


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--minutes', type=float, default=10.0)
    parser.add_argument('--seed', type=int, default=1)
    parser.add_argument('--only', default='')
    args = parser.parse_args()
    names = [fn for fn, weight in CASES for _ in range(weight) if not args.only or args.only in fn.__name__]
    deadline = time.time() + 60 * args.minutes
    done = 0
    while time.time() < deadline:
        seed = args.seed * 1_000_003 + done
        rng = np.random.default_rng(seed)
        fn = names[int(rng.integers(0, len(names)))]
        t0 = time.time()
        try:
            what = fn(rng)
        except AssertionError as error:
            print(f'MISMATCH seed={seed} {fn.__name__}: {error}', flush=True)
            raise
        print(f'[{done}] seed={seed} {what} ({time.time() - t0:.1f}s)', flush=True)
        done += 1
    absent = nat.lib.ds_emu_reads_of_absent_lanes
    absent.restype = __import__('ctypes').c_uint64
    print(f'fuzz: {done} cases, 0 mismatches, {absent()} shuffle reads of absent lanes', flush=True)


if __name__ == '__main__':
    main()
