"""Canonical (hash-seed independent) trigram encoding of title collections - the vectorised host side
of the index build (reference: match_maker.py:84-181, common.py:145-151).

Column ids are the ranks of the distinct trigrams in lexicographic code order and every title's columns
are ascending, so the result does not depend on PYTHONHASHSEED.  The reference's own order (python set
iteration) is reproduced by `MatchMaker.__init__`; this module is the order used for synthetic data,
for the benchmarks and wherever bit-parity with one particular reference process is not the question.
The arithmetic (document frequencies over per-title trigram SETS, idf = math.log(N / df), query-only
trigrams weighted with the maximum idf) is the reference's.
"""
import math

import numpy as np

_ALPHABET = ' abcdefghijklmnopqrstuvwxyz0123456789'
_BASE = len(_ALPHABET)            # 37 -> at most 50,653 trigrams, fits u16 column ids
_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate(_ALPHABET):
    _LUT[ord(_ch)] = _i


def title_trigram_sets(titles, chunk=200000):
    """CSR (ptr int64[n+1], codes int32[nnz]) of each title's DISTINCT trigram codes, ascending.
    Follows common.get_n_grams (common.py:150-151): every 3-character window, as a set."""
    ptr = np.zeros(len(titles) + 1, dtype=np.int64)
    parts = []
    for c0 in range(0, len(titles), chunk):
        block = titles[c0:c0 + chunk]
        lengths = np.fromiter((len(t) for t in block), dtype=np.int64, count=len(block))
        raw = np.frombuffer(''.join(block).encode('latin-1', 'replace'), dtype=np.uint8)
        codes = _LUT[raw]
        if codes.size and codes.max() == 255:
            raise ValueError(f'character {chr(int(raw[np.argmax(codes == 255)]))!r} is outside the title alphabet')
        width = int(lengths.max()) if lengths.size else 0
        n_tri = max(width - 2, 0)
        offsets = np.zeros(len(block) + 1, dtype=np.int64)
        np.cumsum(lengths, out=offsets[1:])
        padded = np.zeros((len(block), width + 2), dtype=np.int32)
        row_of = np.repeat(np.arange(len(block)), lengths)
        col_of = np.arange(codes.size) - np.repeat(offsets[:-1], lengths)
        padded[row_of, col_of] = codes
        tri = (padded[:, :n_tri] * _BASE + padded[:, 1:n_tri + 1]) * _BASE + padded[:, 2:n_tri + 2]
        sentinel = _BASE ** 3
        valid = np.arange(n_tri)[None, :] < (lengths - 2)[:, None]
        tri = np.where(valid, tri, sentinel)
        tri.sort(axis=1)
        keep = tri != sentinel
        keep[:, 1:] &= tri[:, 1:] != tri[:, :-1]
        counts = keep.sum(axis=1)
        np.cumsum(counts, out=ptr[c0 + 1:c0 + 1 + len(block)])
        ptr[c0 + 1:c0 + 1 + len(block)] += ptr[c0]
        parts.append(tri[keep].astype(np.int32))
    codes = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int32)
    return ptr, codes


def encode_canonical(test_titles, truth_titles):
    """-> dict(idf64 f64[V], t_ptr, t_cols u16, q_ptr, q_cols u16, vocab_codes int32[V], n_truth)."""
    t_ptr, t_codes = title_trigram_sets(truth_titles)
    q_ptr, q_codes = title_trigram_sets(test_titles)
    vocab = np.union1d(t_codes, q_codes).astype(np.int32)
    if vocab.size > 65535:
        raise ValueError(f'{vocab.size} distinct trigrams: more than the 65535 u16 column ids')
    t_cols = np.searchsorted(vocab, t_codes).astype(np.uint16)
    q_cols = np.searchsorted(vocab, q_codes).astype(np.uint16)
    n_truth = len(truth_titles)
    df = np.bincount(t_cols, minlength=vocab.size)                      # sets => one count per title
    idf64 = np.zeros(vocab.size, dtype=np.float64)
    in_truth = df > 0
    idf64[in_truth] = [math.log(n_truth / int(c)) for c in df[in_truth]]  # match_maker.py:139 (math.log)
    max_idf = float(idf64[in_truth].max()) if in_truth.any() else 0.0    # :95
    idf64[~in_truth] = max_idf                                           # :151, :180-181
    return dict(idf64=idf64, t_ptr=t_ptr, t_cols=t_cols, q_ptr=q_ptr, q_cols=q_cols, vocab_codes=vocab, n_truth=n_truth)


def trigram_text(code):
    c0, rest = divmod(int(code), _BASE * _BASE)
    c1, c2 = divmod(rest, _BASE)
    return _ALPHABET[c0] + _ALPHABET[c1] + _ALPHABET[c2]
