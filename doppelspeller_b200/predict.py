"""Batched form of the fuzzy pre-match cascade of `Prediction`
(reference: /root/reference/doppelspeller/predict.py:140-183).

    Prediction._get_levenshtein_deletion_ratio   predict.py:140-145
    Prediction._get_levenshtein_ratio            predict.py:147-156
    the `> 94`, group-max and ambiguity filter   predict.py:158-176

The reference maps a Python lambda over every (title, candidate) pair.  `get_levenshtein_ratios` takes the same two
lists of strings (host filter, two kernel calls); `get_levenshtein_ratios_indexed` + `select_close_matches_grouped` are
the device-resident form (title tables + pair indexes in, ratios / chosen pair per title out, no host round trip).
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


class PrematchTables:
    """Device-resident string tables of one title collection for the indexed pre-match: the raw titles and
    their token-sorted forms (`' '.join(sorted(title.split()))`, common.py:166 - sorted once per TITLE, not once
    per pair like the reference's per-pair lambda does)."""

    def __init__(self, titles, device=0):
        import torch

        from .common import _string_table, _token_sort
        self.device = torch.device('cuda', device)
        raw, raw_off = _string_table(titles)
        srt, srt_off = _string_table([_token_sort(t) for t in titles])
        to_dev = lambda x: torch.as_tensor(x).to(self.device)   # noqa: E731
        self.raw, self.raw_off, self.sorted, self.sorted_off = to_dev(raw), to_dev(raw_off), to_dev(srt), to_dev(srt_off)
        self.lengths = to_dev(np.diff(raw_off))
        self.n = len(titles)


def _ratio_pairs(bytes_a, off_a, n_a, bytes_b, off_b, n_b, idx_a, idx_b):
    import torch

    from . import _native as nat
    n = int(idx_a.shape[0])
    out = torch.empty(n, dtype=torch.int32, device=idx_a.device)
    if n:
        nat.check(nat.lib.ds_levenshtein_ratio_pairs(nat.ptr(bytes_a), nat.ptr(off_a), n_a, nat.ptr(bytes_b), nat.ptr(off_b), n_b,
                                                     nat.ptr(idx_a), nat.ptr(idx_b), n, nat.ptr(out), nat.stream_for(idx_a)))
    return out


def get_levenshtein_ratios_indexed(tables_a, tables_b, idx_a, idx_b, threshold=LEVENSHTEIN_RATIO_THRESHOLD):
    """Prediction._get_levenshtein_ratio (predict.py:140-156) for the pairs (tables_a[idx_a[p]], tables_b[idx_b[p]]),
    entirely on the GPU and without a host synchronisation (ds_prematch_pairs): float64 length pre-filter,
    levenshtein_ratio of the survivors, token-sort ratio of those at or below the threshold.  idx_* are int32 CUDA
    tensors; returns an int32 CUDA tensor."""
    import torch

    from . import _native as nat
    idx_a, idx_b = idx_a.to(torch.int32).contiguous(), idx_b.to(torch.int32).contiguous()
    n = int(idx_a.shape[0])
    out = torch.empty(n, dtype=torch.int32, device=idx_a.device)
    if n:
        nat.check(nat.lib.ds_prematch_pairs(
            nat.ptr(tables_a.raw), nat.ptr(tables_a.raw_off), nat.ptr(tables_a.sorted), nat.ptr(tables_a.sorted_off), tables_a.n,
            nat.ptr(tables_b.raw), nat.ptr(tables_b.raw_off), nat.ptr(tables_b.sorted), nat.ptr(tables_b.sorted_off), tables_b.n,
            nat.ptr(idx_a), nat.ptr(idx_b), n, int(threshold), nat.ptr(out), nat.stream_for(idx_a)))
    return out


def select_close_matches_grouped(ratios, n_titles, run, invalid=None, threshold=LEVENSHTEIN_RATIO_THRESHOLD):
    """predict.py:158-176 for a [n_titles, run] candidate list (ds_select_close_matches): int64[n_titles], the index of
    the pair kept as the title's "very close" match or -1.  `ratios` int32 (CUDA tensor: stays on the device; numpy:
    staged), `invalid` optional uint8 mask of pairs to leave out."""
    from . import _native as nat
    cuda = hasattr(ratios, 'is_cuda') and ratios.is_cuda
    if cuda:
        import torch
        out = torch.empty(n_titles, dtype=torch.int64, device=ratios.device)
        ratios = ratios.contiguous()
    else:
        out = np.empty(n_titles, dtype=np.int64)
        ratios = np.ascontiguousarray(ratios, dtype=np.int32)
    if int(ratios.shape[0]) != n_titles * run:
        raise ValueError('ratios must hold n_titles * run entries')
    nat.check(nat.lib.ds_select_close_matches(nat.ptr(ratios), nat.ptr(invalid), n_titles, int(run), int(threshold), nat.ptr(out),
                                              nat.stream_for(ratios)))
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
