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


def title_table(titles):
    """Compact (bytes uint8[total], offsets int64[n+1]) table of ASCII titles (truncated to 255 like transform_title)."""
    lengths = np.fromiter(map(len, titles), dtype=np.int64, count=len(titles))       # C-level iteration: no generator frame per title
    if lengths.size and lengths.max() > 255:
        titles = [t[:255] for t in titles]
        np.minimum(lengths, 255, out=lengths)
    offsets = np.zeros(len(titles) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    data = np.frombuffer(''.join(titles).encode('latin-1', 'replace'), dtype=np.uint8)
    return (np.array(data) if data.size else np.zeros(1, dtype=np.uint8)), offsets


def encode_canonical_device(test_titles, truth_titles, device=0, truth_table=None, query_table=None):
    """`encode_canonical` on the GPU (csrc/ds_encode.cu): same outputs, as CUDA tensors that plug straight into
    TruthIndex / MatchMaker.from_encoded without touching the host again.  `truth_table` / `query_table` = (bytes,
    offsets) CUDA tensors of a side that is already resident (skips the host string join and the H2D copy)."""
    import ctypes

    import torch

    from . import _native as nat
    dev = torch.device('cuda', device)
    to_dev = lambda x: torch.as_tensor(x).to(dev)   # noqa: E731
    if truth_table is None:
        t_bytes, t_off = title_table(truth_titles)
        d_tb, d_to = to_dev(t_bytes), to_dev(t_off)
        truth_total = int(t_off[-1])
    else:
        d_tb, d_to = truth_table
        truth_total = int(d_tb.shape[0])
    n_truth = int(d_to.shape[0]) - 1
    if query_table is None:
        q_bytes, q_off = title_table(test_titles)
        d_qb, d_qo = to_dev(q_bytes), to_dev(q_off)
        n_queries, query_total = len(test_titles), int(q_off[-1])
    else:
        d_qb, d_qo = query_table
        n_queries, query_total = int(d_qo.shape[0]) - 1, int(d_qb.shape[0])
    max_vocab = int(nat.lib.ds_encode_max_vocab())
    t_ptr = torch.empty(n_truth + 1, dtype=torch.int64, device=dev)
    q_ptr = torch.empty(n_queries + 1, dtype=torch.int64, device=dev)
    t_cols = torch.empty(max(1, truth_total), dtype=torch.uint16, device=dev)
    q_cols = torch.empty(max(1, query_total), dtype=torch.uint16, device=dev)
    idf64 = torch.empty(max_vocab, dtype=torch.float64, device=dev)
    vocab = torch.empty(max_vocab, dtype=torch.int32, device=dev)
    n_vocab, t_nnz, q_nnz = ctypes.c_int32(0), ctypes.c_int64(0), ctypes.c_int64(0)
    with torch.cuda.device(dev):
        nat.check(nat.lib.ds_encode_trigrams(
            nat.ptr(d_tb), nat.ptr(d_to), n_truth, nat.ptr(d_qb), nat.ptr(d_qo), n_queries, nat.ptr(t_ptr),
            nat.ptr(t_cols), nat.ptr(q_ptr), nat.ptr(q_cols), nat.ptr(idf64), nat.ptr(vocab), ctypes.byref(n_vocab),
            ctypes.byref(t_nnz), ctypes.byref(q_nnz), device, torch.cuda.current_stream(dev).cuda_stream))
    return dict(idf64=idf64[:n_vocab.value], t_ptr=t_ptr, t_cols=t_cols[:t_nnz.value], q_ptr=q_ptr, q_cols=q_cols[:q_nnz.value],
                vocab_codes=vocab[:n_vocab.value], n_truth=n_truth)
