"""Drop-in title normalisation and fuzzy ratios (reference: /root/reference/doppelspeller/common.py:20-47,161-167).

`transform_title(title)` / `transform_titles(titles)` run common.transform_title on the GPU
(csrc/ds_encode.cu `k_transform`); the only host work is handing the strings over as UTF-32.

`levenshtein_ratio(text, text_to_match)` = int(round(Levenshtein.ratio(a, b) * 100)) where `ratio` is
python-levenshtein 0.12.0's (la + lb - indel) / (la + lb) (third-party C code that is not part of the
reference tree: restated from its documented definition, parity unpinned - SURVEY.md 8(c)).
The batch forms are what `Prediction` should call; the scalar forms keep the reference signatures.
"""
import logging
import unicodedata

import numpy as np

from . import _native as nat

LOGGER = logging.getLogger(__name__)
N_GRAMS = 3                                   # settings.py:15
MAX_CHARACTERS_ALLOWED_IN_THE_TITLE = 255     # settings.py:68 (uint8 max)
_ASCII_OF_CP = None


def ascii_of_codepoint_table():
    """uint8 table: code point -> the ASCII character its canonical decomposition (NFD) contains, 0 = none.
    `unicodedata.normalize('NFD', x).encode('ascii', 'ignore')` (common.py:25-26) keeps exactly these characters;
    no code point contributes more than one, and none above U+226F contributes any (asserted here)."""
    global _ASCII_OF_CP
    if _ASCII_OF_CP is None:
        table = np.zeros(0x2270, dtype=np.uint8)
        for cp in range(1, 0x110000):
            if 0xD800 <= cp <= 0xDFFF:
                continue
            kept = unicodedata.normalize('NFD', chr(cp)).encode('ascii', 'ignore')
            if kept:
                if len(kept) != 1 or cp >= table.shape[0]:
                    raise RuntimeError(f'unexpected canonical decomposition of U+{cp:04X} in this Unicode database')
                table[cp] = kept[0]
        _ASCII_OF_CP = table
    return _ASCII_OF_CP


def _utf32_table(titles):
    """(code points uint32[total], offsets int64[n+1]) of python strings."""
    lengths = np.fromiter(map(len, titles), dtype=np.int64, count=len(titles))
    offsets = np.zeros(len(titles) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    data = np.frombuffer(''.join(titles).encode('utf-32-le', 'surrogatepass'), dtype=np.uint32)
    return (np.array(data) if data.size else np.zeros(1, dtype=np.uint32)), offsets


def transform_titles_table(titles, device=None, warn=True):
    """common.transform_title over a sequence of raw titles -> (bytes uint8[total], offsets int64[n+1], raw_len
    int32[n]) on the host, or CUDA tensors when `device` is given (they feed encode.encode_canonical_device /
    the ds_*_pairs entry points without touching the host again)."""
    cps, offsets = _utf32_table(titles)
    n = len(titles)
    table = ascii_of_codepoint_table()
    capacity = int(offsets[-1]) + 3 * n + 1
    if device is None:
        out = np.zeros(capacity, dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.int64)
        raw = np.zeros(max(1, n), dtype=np.int32)
        nat.check(nat.lib.ds_transform_titles(nat.ptr(cps), nat.ptr(offsets), n, nat.ptr(table), int(table.shape[0]), nat.ptr(out),
                                              nat.ptr(out_off), nat.ptr(raw), nat.current_device(), nat.current_stream()))
        out = out[:max(1, int(out_off[-1]))]
        raw = raw[:n]
        raw_host = raw
    else:
        import torch
        dev = torch.device('cuda', int(device))
        out = torch.zeros(capacity, dtype=torch.uint8, device=dev)
        out_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        raw = torch.zeros(max(1, n), dtype=torch.int32, device=dev)
        d_cps, d_offsets, d_table = (torch.as_tensor(x).to(dev) for x in (cps, offsets, table))
        with torch.cuda.device(dev):
            nat.check(nat.lib.ds_transform_titles(nat.ptr(d_cps), nat.ptr(d_offsets), n, nat.ptr(d_table), int(table.shape[0]),
                                                  nat.ptr(out), nat.ptr(out_off), nat.ptr(raw), int(device),
                                                  torch.cuda.current_stream(dev).cuda_stream))
        total = int(out_off[-1].item()) if n else 0
        out = out[:max(1, total)]
        raw = raw[:n]
        raw_host = raw.cpu().numpy() if warn else None
    if warn and n:
        for i in np.nonzero(raw_host < N_GRAMS)[0]:                                   # common.py:34-38
            LOGGER.warning(f"Title ({titles[i]}) less than length {N_GRAMS} found, after transforming the title. "
                           f"Pre-pending 0's!\n")
        for i in np.nonzero(raw_host > MAX_CHARACTERS_ALLOWED_IN_THE_TITLE)[0]:       # common.py:40-45
            LOGGER.warning(f'Titles greater than length {MAX_CHARACTERS_ALLOWED_IN_THE_TITLE} are not allowed. '
                           f'Trimming the title ({titles[i][:10]}...)!')
    return out, out_off, raw


def transform_titles(titles):
    """[common.transform_title(t) for t in titles] (common.py:20-47), computed on the GPU."""
    titles = list(titles)
    out, out_off, _ = transform_titles_table(titles)
    blob = out.tobytes().decode('latin-1')
    return [blob[out_off[i]:out_off[i + 1]] for i in range(len(titles))]


def transform_title(title):
    """Transforms a title in to alpha-numeric-only (plus spaces) text (common.py:20-47)."""
    return transform_titles([title])[0]


def _string_table(texts):
    """(bytes uint8[total], offsets int64[n+1]) of latin-1 encodable strings."""
    lengths = np.fromiter(map(len, texts), dtype=np.int64, count=len(texts))      # latin-1: one byte per character
    if lengths.size and lengths.max() > nat.MAX_TITLE:
        raise ValueError('titles longer than 255 characters are outside the domain of the ratio kernels')
    offsets = np.zeros(len(texts) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    data = np.frombuffer(''.join(texts).encode('latin-1'), dtype=np.uint8)
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
