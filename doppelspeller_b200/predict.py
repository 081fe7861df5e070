"""Batched form of the fuzzy pre-match cascade of `Prediction`
(reference: /root/reference/doppelspeller/predict.py:140-183).

    Prediction._get_levenshtein_deletion_ratio   predict.py:140-145
    Prediction._get_levenshtein_ratio            predict.py:147-156
    the `> 94`, group-max and ambiguity filter   predict.py:158-176

The reference maps a Python lambda over every (title, candidate) pair; here the length pre-filter is one
vectorised numpy expression and the two ratios run as two kernel launches over the surviving pairs.
"""
import numpy as np

from .common import levenshtein_ratio_batch, levenshtein_token_sort_ratio_batch

LEVENSHTEIN_RATIO_THRESHOLD = 94      # settings.py:75


def get_levenshtein_deletion_ratios(lengths_x, lengths_y):
    """predict.py:140-145, float64, same association: ((total - delta) / total) * 100."""
    lengths_x = np.asarray(lengths_x, dtype=np.int64)
    lengths_y = np.asarray(lengths_y, dtype=np.int64)
    total = (lengths_x + lengths_y).astype(np.float64)
    delta = np.abs(lengths_x - lengths_y).astype(np.float64)
    with np.errstate(all='ignore'):
        return ((total - delta) / total) * 100


def get_levenshtein_ratios(titles, matches, threshold=LEVENSHTEIN_RATIO_THRESHOLD):
    """Element-wise Prediction._get_levenshtein_ratio (predict.py:147-156) -> int32 array."""
    n = len(titles)
    out = np.zeros(n, dtype=np.int32)
    if n == 0:
        return out
    lengths_x = np.fromiter((len(t) for t in titles), dtype=np.int64, count=n)
    lengths_y = np.fromiter((len(t) for t in matches), dtype=np.int64, count=n)
    keep = np.nonzero(~(get_levenshtein_deletion_ratios(lengths_x, lengths_y) < threshold))[0]
    if keep.size == 0:
        return out
    kept_x = [titles[i] for i in keep]
    kept_y = [matches[i] for i in keep]
    ratios = levenshtein_ratio_batch(kept_x, kept_y)
    again = np.nonzero(ratios <= threshold)[0]
    if again.size:
        ratios[again] = levenshtein_token_sort_ratio_batch([kept_x[i] for i in again], [kept_y[i] for i in again])
    out[keep] = ratios
    return out


def select_close_matches(test_index, ratios, threshold=LEVENSHTEIN_RATIO_THRESHOLD):
    """predict.py:158-176: positions of the pairs kept as "very close" matches: ratio > threshold, equal to
    the maximum of their test_index, and that maximum attained exactly once."""
    test_index = np.asarray(test_index)
    ratios = np.asarray(ratios)
    candidates = np.nonzero(ratios > threshold)[0]
    if candidates.size == 0:
        return candidates
    keys = test_index[candidates]
    order = np.lexsort((-ratios[candidates], keys))
    keys_sorted, ratios_sorted = keys[order], ratios[candidates][order]
    first = np.ones(keys_sorted.size, dtype=bool)
    first[1:] = keys_sorted[1:] != keys_sorted[:-1]
    group_start = np.maximum.accumulate(np.where(first, np.arange(keys_sorted.size), 0))
    is_max = ratios_sorted == ratios_sorted[group_start]
    n_max = np.bincount(np.cumsum(first) - 1, weights=is_max).astype(np.int64)
    unique = n_max[np.cumsum(first) - 1] == 1
    return np.sort(candidates[order][is_max & unique])
