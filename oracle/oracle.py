"""CPU oracle for the DoppelSpeller hot path - TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module.  The product package (`doppelspeller_b200`) never does: it fails loudly when its
CUDA library is missing instead of falling back to anything here.

What it is: a ctypes front-end to `oracle/ds_oracle.c` (plain C restatement of the reference's numba
kernels, each function citing reference file:line) plus numpy/Python restatements of the host-side
index construction (`MatchMaker.__init__`, match_maker.py:84-181) and title encoding
(feature_engineering.py:298-319).

Parity pinning: `tests/test_oracle_vs_reference.py` compares every function here with the
reference's own jitted functions imported from /root/reference (container only), and
`tests/golden/*.npz` (minted by `tests/golden/make_golden.py` from the reference) pin it on the GPU
box.  `levenshtein_ratio` / token-sort follow third-party python-levenshtein==0.12.0
(requirements.txt:9, source not in the reference tree): parity UNPINNED for those two.
"""
import ctypes
import math
import os
import subprocess
from collections import Counter

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'libds_oracle.so')
_LIB = None

N_WORDS = 15
N_FEATURES = 66
MAX_CHARS = 255
ALPHABET = '- abcdefghijklmnopqrstuvwxyz0123456789'   # feature_engineering.py:200
SPACE_CODE = 1

_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_u8p = ctypes.POINTER(ctypes.c_uint8)
_c_u32p = ctypes.POINTER(ctypes.c_uint32)
_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compiles oracle/ds_oracle.c with the committed Makefile (gcc, no fast-math)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, 'ds_oracle.c')):
        subprocess.run(['make', '-C', _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(_LIB_PATH)
        _LIB.orc_fast_arg_top_k.restype = ctypes.c_int64
        _LIB.orc_indel_ratio_u8.restype = ctypes.c_int
        _LIB.orc_indel_distance_u8.restype = ctypes.c_int
        _LIB.orc_indel_distance.restype = ctypes.c_int
        _LIB.orc_levenshtein_ratio.restype = ctypes.c_int
        _LIB.orc_prefilter_rejects.restype = ctypes.c_int
        _LIB.orc_topn_batch.restype = ctypes.c_int
        _LIB.orc_max_threads.restype = ctypes.c_int
        _LIB.orc_py_float_sum_mode.restype = ctypes.c_double
    return _LIB


def _p(array, ctype):
    return array.ctypes.data_as(ctypes.POINTER(ctype))


def max_threads():
    return int(lib().orc_max_threads())


# ----------------------------------------------------------------------------------------------------
# Index construction (host side of MatchMaker) - match_maker.py:84-181, common.py:145-151
# ----------------------------------------------------------------------------------------------------
def get_n_grams(title, n=3):
    """common.py:150-151"""
    return set([title[i:i + n] for i in range(len(title)) if len(title[i:i + n]) == n])


def encode_reference_order(data_n_grams, truth_n_grams):
    """Restates MatchMaker.__init__ (match_maker.py:84-109) on two sequences of python n-gram SETS.

    Column ids and the per-truth-row accumulation order depend on python set iteration order
    (match_maker.py:144-147, :172-174), so this must run in the same process (same PYTHONHASHSEED) as the
    reference instance it is compared with, on the same set objects.

    Returns a dict of numpy arrays (the "encoded index"):
      vocab      list[str]   n-gram of every column id                       (:144-147)
      w64, w32   [V]         idf (float64) / float32(idf); query-only n-grams get max idf (:151, :180-181)
      t_ptr/t_cols           truth rows CSR, columns in SET-ITERATION order  (:167-175)
      q_ptr/q_cols           query rows CSR, columns in set-iteration order
    """
    counter = Counter(x for y in data_n_grams for x in set(y))                    # common.py:145-147
    counter_truth = Counter(x for y in truth_n_grams for x in set(y))
    n_truth = len(truth_n_grams)
    idf = {key: math.log(n_truth / count) for key, count in counter_truth.items()}   # :135-142
    max_idf = max(idf.values())                                                      # :95
    all_n_grams = set(list(counter.keys()) + list(counter_truth.keys()))             # :144-147
    vocab = list(all_n_grams)
    encoding = {g: i for i, g in enumerate(vocab)}
    w64 = np.array([idf.get(g, max_idf) for g in vocab], dtype=np.float64)
    w32 = w64.astype(np.float32)

    def to_csr(rows):
        ptr = np.zeros(len(rows) + 1, dtype=np.int64)
        cols = []
        for i, value in enumerate(rows):
            cols.extend(encoding[x] for x in value)
            ptr[i + 1] = len(cols)
        return ptr, np.array(cols, dtype=np.int32)

    t_ptr, t_cols = to_csr(truth_n_grams)
    q_ptr, q_cols = to_csr(data_n_grams)
    return dict(vocab=vocab, n_truth=n_truth, w64=w64, w32=w32, t_ptr=t_ptr, t_cols=t_cols, q_ptr=q_ptr,
                q_cols=q_cols)


def finish_index(enc, queries=None):
    """From an encoded index (column ids + per-row order given) derive what fast_jaccard consumes
    (`queries`: only the posting lists these query rows touch are built - the others stay empty; for parity samples
    over very large truth sets, where sorting every posting would take minutes):
      sums       f32 sequential sum per truth row in the given order          (match_maker.py:172-174)
      post_ptr/post_rows  per column ascending truth rows, zero weights dropped (lil semantics, :122-133)
      qs_ptr/qs_cols      per query ASCENDING column ids with w32 != 0          (:111-120)
      q_mx       None: python's sum() of w64 over qs_cols (:197) is restated in C (orc_py_float_sum)
    """
    L = lib()
    n_truth = int(enc['n_truth'])
    w32 = np.ascontiguousarray(enc['w32'], dtype=np.float32)
    w64 = np.ascontiguousarray(enc['w64'], dtype=np.float64)
    n_vocab = w32.shape[0]
    t_ptr = np.ascontiguousarray(enc['t_ptr'], dtype=np.int64)
    t_cols = np.ascontiguousarray(enc['t_cols'], dtype=np.int32)
    sums = np.zeros(n_truth, dtype=np.float32)
    L.orc_truth_sums(ctypes.c_int64(n_truth), _p(t_ptr, ctypes.c_int64), _p(t_cols, ctypes.c_int32),
                     _p(w32, ctypes.c_float), _p(sums, ctypes.c_float))
    # postings: (col, row) sorted by col then row, dropping zero weights
    keep = w32[t_cols] != 0
    if queries is not None:
        q_ptr_all = np.asarray(enc['q_ptr'], dtype=np.int64)
        q_cols_all = np.asarray(enc['q_cols'])
        needed = np.zeros(n_vocab, dtype=bool)
        for q in np.asarray(queries, dtype=np.int64):
            needed[q_cols_all[q_ptr_all[q]:q_ptr_all[q + 1]]] = True
        keep &= needed[t_cols]
    kept = np.nonzero(keep)[0]
    cols_k = t_cols[kept].astype(np.int64)
    rows_k = (np.searchsorted(t_ptr, kept, side='right') - 1).astype(np.int64)
    order = np.argsort((cols_k << 32) | rows_k, kind='stable')       # by column, then by row
    post_rows = rows_k[order].astype(np.int32)
    post_ptr = np.zeros(n_vocab + 1, dtype=np.int64)
    np.cumsum(np.bincount(cols_k, minlength=n_vocab), out=post_ptr[1:])
    # queries: ascending, zero weights dropped
    q_ptr = np.ascontiguousarray(enc['q_ptr'], dtype=np.int64)
    q_cols = np.ascontiguousarray(enc['q_cols'], dtype=np.int32)
    n_q = q_ptr.shape[0] - 1
    q_rows = np.repeat(np.arange(n_q, dtype=np.int64), np.diff(q_ptr))
    keep_q = w32[q_cols] != 0
    qc, qr = q_cols[keep_q].astype(np.int64), q_rows[keep_q]
    order_q = np.lexsort((qc, qr))
    qs_cols = qc[order_q].astype(np.int32)
    qs_ptr = np.zeros(n_q + 1, dtype=np.int64)
    np.cumsum(np.bincount(qr, minlength=n_q), out=qs_ptr[1:])
    out = dict(enc)
    out.update(sums=sums, post_ptr=post_ptr, post_rows=post_rows, qs_ptr=qs_ptr, qs_cols=qs_cols,
               q_mx=None)   # NULL -> orc_topn_batch restates python's sum() (:197)
    return out


def py_float_sum(w64, cols, compensated=True):
    """C restatement of `sum([w64[c] for c in cols])` as CPython >= 3.12 executes it (Neumaier)."""
    w64 = np.ascontiguousarray(w64, dtype=np.float64)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    return float(lib().orc_py_float_sum_mode(_p(w64, ctypes.c_double), _p(cols, ctypes.c_int32),
                                             ctypes.c_int64(cols.shape[0]), ctypes.c_int(int(compensated))))


def fast_jaccard(index, q):
    """match_maker.py:16-50 for query row q of a finished index -> float64[N]."""
    L = lib()
    n_truth = int(index['n_truth'])
    cols = np.ascontiguousarray(index['qs_cols'][index['qs_ptr'][q]:index['qs_ptr'][q + 1]])
    if index.get('q_mx') is not None:
        mx = float(index['q_mx'][q])
    else:
        mx = float(sum([float(index['w64'][col]) for col in cols]))   # the interpreter's own sum(), like :197
    out = np.zeros(n_truth, dtype=np.float64)
    L.orc_fast_jaccard(ctypes.c_int64(n_truth), ctypes.c_double(mx), _p(cols, ctypes.c_int32),
                       ctypes.c_int64(cols.shape[0]), _p(index['post_ptr'], ctypes.c_int64),
                       _p(index['post_rows'], ctypes.c_int32), _p(index['w32'], ctypes.c_float),
                       _p(index['sums'], ctypes.c_float), _p(out, ctypes.c_double))
    return out


def fast_arg_top_k(array, k):
    """match_maker.py:53-71 -> int64 rows, descending index, at most k."""
    array = np.ascontiguousarray(array, dtype=np.float64)
    out = np.zeros(max(k, 1), dtype=np.int64)
    kth = ctypes.c_float(0)
    count = lib().orc_fast_arg_top_k(_p(array, ctypes.c_double), ctypes.c_int64(array.shape[0]),
                                     ctypes.c_int64(k), _p(out, ctypes.c_int64), ctypes.byref(kth))
    return out[:count]


def topn(index, k, queries=None, n_threads=0):
    """Batch of get_closest_matches (match_maker.py:192-203) -> (rows int64[Q,k] descending (-1 pad),
    count int64[Q], kth_key f32[Q])."""
    L = lib()
    qs_ptr, qs_cols = index['qs_ptr'], index['qs_cols']
    if queries is not None:
        queries = np.asarray(queries, dtype=np.int64)
        lens = (qs_ptr[queries + 1] - qs_ptr[queries])
        new_ptr = np.zeros(len(queries) + 1, dtype=np.int64)
        np.cumsum(lens, out=new_ptr[1:])
        parts = [qs_cols[qs_ptr[q]:qs_ptr[q + 1]] for q in queries]
        qs_cols = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0, np.int32), dtype=np.int32)
        q_mx = index['q_mx'][queries] if index.get('q_mx') is not None else None
        qs_ptr = new_ptr
    else:
        q_mx = index.get('q_mx')
    n_q = qs_ptr.shape[0] - 1
    rows = np.full((n_q, k), -1, dtype=np.int64)
    count = np.zeros(n_q, dtype=np.int64)
    kth = np.zeros(n_q, dtype=np.float32)
    q_mx_p = _p(np.ascontiguousarray(q_mx, dtype=np.float64), ctypes.c_double) if q_mx is not None else None
    rc = L.orc_topn_batch(ctypes.c_int64(int(index['n_truth'])), _p(index['post_ptr'], ctypes.c_int64),
                          _p(index['post_rows'], ctypes.c_int32), _p(index['w32'], ctypes.c_float),
                          _p(index['w64'], ctypes.c_double), _p(index['sums'], ctypes.c_float),
                          ctypes.c_int64(n_q), _p(qs_ptr, ctypes.c_int64), _p(qs_cols, ctypes.c_int32),
                          q_mx_p, ctypes.c_int64(k), _p(rows, ctypes.c_int64), _p(count, ctypes.c_int64),
                          _p(kth, ctypes.c_float), ctypes.c_int(n_threads))
    if rc != 0:
        raise MemoryError('orc_topn_batch failed')
    return rows, count, kth


# ----------------------------------------------------------------------------------------------------
# Title encoding - feature_engineering.py:298-319
# ----------------------------------------------------------------------------------------------------
_ENCODING = {ch: i for i, ch in enumerate(ALPHABET)}


def encode_title(title):
    """feature_engineering.py:298-307: codes, zero padded / truncated to 255."""
    out = np.zeros(MAX_CHARS, dtype=np.uint8)
    codes = [_ENCODING[ch] for ch in title[:MAX_CHARS]]
    out[:len(codes)] = codes
    return out


def truth_words_counts(title, words_counter):
    """feature_engineering.py:309-319: first 15 words' document frequencies, zero padded."""
    out = np.zeros(N_WORDS, dtype=np.uint32)
    counts = [words_counter.get(w) for w in title.split()][:N_WORDS]
    out[:len(counts)] = counts
    return out


# ----------------------------------------------------------------------------------------------------
# Pair scoring
# ----------------------------------------------------------------------------------------------------
def indel_ratio_u8(a, b):
    """fast_levenshtein_ratio (feature_engineering.py:25-63) on two code arrays -> int 0..255."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    return int(lib().orc_indel_ratio_u8(_p(a, ctypes.c_uint8), ctypes.c_int(a.shape[0]),
                                        _p(b, ctypes.c_uint8), ctypes.c_int(b.shape[0])))


def indel_ratio_u8_batch(a, b, la, lb, n_threads=0):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    la = np.ascontiguousarray(la, dtype=np.uint8)
    lb = np.ascontiguousarray(lb, dtype=np.uint8)
    n = la.shape[0]
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_indel_ratio_u8_batch(_p(a, ctypes.c_uint8), _p(b, ctypes.c_uint8), ctypes.c_int64(a.shape[1]),
                                   _p(la, ctypes.c_uint8), _p(lb, ctypes.c_uint8), ctypes.c_int64(n),
                                   _p(out, ctypes.c_uint8), ctypes.c_int(n_threads))
    return out


def construct_features(la, lb, title, truth, counts, space_code, n_truth, n_threads=0):
    """construct_features gufunc (feature_engineering.py:75-169) over P pairs -> float32[P,66]."""
    la = np.ascontiguousarray(la, dtype=np.uint8)
    lb = np.ascontiguousarray(lb, dtype=np.uint8)
    title = np.ascontiguousarray(title, dtype=np.uint8)
    truth = np.ascontiguousarray(truth, dtype=np.uint8)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    n = la.shape[0]
    assert title.shape == truth.shape and title.shape[0] == n and counts.shape == (n, N_WORDS)
    out = np.zeros((n, N_FEATURES), dtype=np.float32)
    lib().orc_construct_features_batch(
        _p(la, ctypes.c_uint8), _p(lb, ctypes.c_uint8), _p(title, ctypes.c_uint8), _p(truth, ctypes.c_uint8),
        ctypes.c_int64(title.shape[1]), _p(counts, ctypes.c_uint32), ctypes.c_uint8(space_code),
        ctypes.c_uint32(n_truth), ctypes.c_int64(n), _p(out, ctypes.c_float), ctypes.c_int(n_threads))
    return out


def _bytes(text):
    return np.frombuffer(text.encode('latin-1', 'replace'), dtype=np.uint8).copy()


def levenshtein_ratio(text, text_to_match):
    """common.py:161-162 with python-levenshtein's ratio restated (parity unpinned)."""
    a, b = _bytes(text), _bytes(text_to_match)
    return int(lib().orc_levenshtein_ratio(_p(a, ctypes.c_uint8), ctypes.c_int(a.shape[0]),
                                           _p(b, ctypes.c_uint8), ctypes.c_int(b.shape[0])))


def levenshtein_token_sort_ratio(text, text_to_match):
    """common.py:165-167"""
    text, text_to_match = ' '.join(sorted(text.split())), ' '.join(sorted(text_to_match.split()))
    return levenshtein_ratio(text, text_to_match)


def prematch_ratio(x, y, threshold=94):
    """Prediction._get_levenshtein_ratio (predict.py:140-156)."""
    if lib().orc_prefilter_rejects(ctypes.c_int(len(x)), ctypes.c_int(len(y))):
        return 0
    ratio = levenshtein_ratio(x, y)
    if ratio <= threshold:
        return levenshtein_token_sort_ratio(x, y)
    return ratio


# ---------------------------------------------------------------------------------------------------
# f3: transform_title (common.py:20-47), restated character by character (no regular expressions) so that
# it is an independent check of both the reference's regex formulation and the CUDA state machine
# ---------------------------------------------------------------------------------------------------
def transform_title(title, n_grams=3, max_chars=255):
    import unicodedata
    kept = []
    for ch in unicodedata.normalize('NFD', title):                      # common.py:25
        if ord(ch) >= 128:                                              # .encode('ascii', 'ignore')  :26
            continue
        ch = ch.lower()                                                 # .lower()                    :26
        if ch == '-':                                                   # .replace('-', ' ')          :26
            ch = ' '
        if ch.isalnum() or ch.isspace():                                # KEEP_REGEX [a-zA-Z0-9\s]    :28
            kept.append(ch)
    collapsed = []
    for ch in kept:                                                     # SUBSTITUTE_REGEX ' +' -> ' ' :30
        if ch == ' ' and collapsed and collapsed[-1] == ' ':
            continue
        collapsed.append(ch)
    text = ''.join(collapsed).strip()                                   # .strip()                    :30
    number_of_characters = len(text)                                    # :31
    text = text[:max_chars].strip()                                     # :32
    if number_of_characters < n_grams:                                  # :34-38
        return text.rjust(n_grams, '0')
    return text


# ---------------------------------------------------------------------------------------------------
# f4: gradient-boosted tree inference.  xgboost==0.90 (requirements.txt:8) is a third-party dependency absent from
# /root/reference; its published CPU prediction path is restated (src/predictor/cpu_predictor.cc PredValue:
# psum = 0.0f; psum += leaf of every tree in order; include/xgboost/tree_model.h GetNext: missing -> default child,
# else fvalue < split_cond ? left : right; src/objective/regression_loss.h: sigmoid in float32).  Parity unpinned.
# ---------------------------------------------------------------------------------------------------
def gbdt_predict(features, nodes, tree_offsets, base_margin=0.0, logistic=True):
    """features float32 [n, f]; nodes: structured array with fields feature/value/yes/no/missing (children relative to
    their tree's first node); returns float32 [n]."""
    x = np.ascontiguousarray(features, dtype=np.float32)
    n = x.shape[0]
    psum = np.zeros(n, dtype=np.float32)
    rows = np.arange(n)
    for t in range(len(tree_offsets) - 1):
        tree = nodes[tree_offsets[t]:tree_offsets[t + 1]]
        at = np.zeros(n, dtype=np.int64)
        while True:
            feature = tree['feature'][at]
            active = feature >= 0
            if not active.any():
                break
            value = x[rows, np.where(active, feature, 0)]
            with np.errstate(invalid='ignore'):
                go = np.where(np.isnan(value), tree['missing'][at], np.where(value < tree['value'][at], tree['yes'][at], tree['no'][at]))
            at = np.where(active, go, at)
        psum = (psum + tree['value'][at]).astype(np.float32)
    margin = (np.float32(base_margin) + psum).astype(np.float32)
    if not logistic:
        return margin
    with np.errstate(over='ignore'):
        return (np.float32(1.0) / (np.float32(1.0) + np.exp(-margin, dtype=np.float32))).astype(np.float32)
