"""Drop-in fuzzy ratios (reference: /root/reference/doppelspeller/common.py:161-167).

`levenshtein_ratio(text, text_to_match)` = int(round(Levenshtein.ratio(a, b) * 100)) where `ratio` is
python-levenshtein 0.12.0's (la + lb - indel) / (la + lb) (third-party C code that is not part of the
reference tree: restated from its documented definition, parity unpinned - SURVEY.md 8(c)).
The batch forms are what `Prediction` should call; the scalar forms keep the reference signatures.
"""
import numpy as np

from . import _native as nat


def _string_table(texts):
    """(bytes uint8[total], offsets int64[n+1]) of latin-1 encodable strings."""
    encoded = [t.encode('latin-1') for t in texts]
    lengths = np.fromiter((len(e) for e in encoded), dtype=np.int64, count=len(encoded))
    if lengths.size and lengths.max() > nat.MAX_TITLE:
        raise ValueError('titles longer than 255 characters are outside the domain of the ratio kernels')
    offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    data = np.frombuffer(b''.join(encoded), dtype=np.uint8)
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    return np.array(data), offsets      # a writable copy (np.frombuffer views are read-only)


def levenshtein_ratio_batch(texts, texts_to_match):
    """Element-wise common.levenshtein_ratio over two equally long sequences of str -> int32 array."""
    if len(texts) != len(texts_to_match):
        raise ValueError('both sequences must have the same length')
    n = len(texts)
    out = np.empty(n, dtype=np.int32)
    if n == 0:
        return out
    bytes_a, off_a = _string_table(texts)
    bytes_b, off_b = _string_table(texts_to_match)
    idx = np.arange(n, dtype=np.int32)
    nat.check(nat.lib.ds_levenshtein_ratio_pairs(nat.ptr(bytes_a), nat.ptr(off_a), n, nat.ptr(bytes_b), nat.ptr(off_b), n,
                                                 nat.ptr(idx), nat.ptr(idx), n, nat.ptr(out), nat.current_stream()))
    return out


def levenshtein_ratio(text, text_to_match):
    """common.py:161-162"""
    return int(levenshtein_ratio_batch([text], [text_to_match])[0])


def _token_sort(text):
    return ' '.join(sorted(text.split()))


def levenshtein_token_sort_ratio_batch(texts, texts_to_match):
    return levenshtein_ratio_batch([_token_sort(t) for t in texts], [_token_sort(t) for t in texts_to_match])


def levenshtein_token_sort_ratio(text, text_to_match):
    """common.py:165-167"""
    return int(levenshtein_token_sort_ratio_batch([text], [text_to_match])[0])
