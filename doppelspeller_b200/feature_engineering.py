"""Drop-in pair scoring (reference: /root/reference/doppelspeller/feature_engineering.py:25-169).

`construct_features` keeps the gufunc call signature of the reference
    construct_features(title_number_of_characters, truth_number_of_characters, title, title_truth,
                       truth_words_counts, space_code, number_of_truth_titles, dummy, response)
(layout '(),(),(l),(l),(m),(),(),(n)->(n)', feature_engineering.py:75-80) and fills `response`;
`fast_levenshtein_ratio(a, b)` keeps the scalar signature.  Both run on the GPU (csrc/ds_pairs.cu).
numpy inputs are host buffers (staged by the library); CUDA torch tensors are used in place.
"""
import numpy as np

from . import _native as nat

NUMBER_OF_WORDS_FEATURES = nat.N_WORDS          # settings.py:65
FEATURES_COUNT = 6 + (4 * NUMBER_OF_WORDS_FEATURES)   # feature_engineering.py:67
MAX_CHARACTERS_ALLOWED_IN_THE_TITLE = nat.MAX_TITLE   # settings.py:68
ALLOWED_CHARACTERS = '- abcdefghijklmnopqrstuvwxyz0123456789'   # feature_engineering.py:200
ENCODING = {character: index for index, character in enumerate(ALLOWED_CHARACTERS)}
SPACE_CODE = ENCODING[' ']


def _is_cuda(x):
    return hasattr(x, 'is_cuda') and x.is_cuda


def _as_2d_u8(x, name):
    nat.expect(x, 'uint8', name)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if x.ndim != 2:
        raise ValueError(f'{name} must be [P, l]')
    return x


def construct_features(title_number_of_characters, truth_number_of_characters, title, title_truth, truth_words_counts,
                       space_code, number_of_truth_titles, dummy=None, response=None):
    """feature_engineering.py:75-169 for P pairs.  Returns `response` (float32 [P, 66])."""
    title = _as_2d_u8(title, 'title')
    title_truth = _as_2d_u8(title_truth, 'title_truth')
    n_pairs, stride = title.shape
    if title_truth.shape != title.shape:
        raise ValueError('title and title_truth must have the same [P, l] shape')
    cuda = _is_cuda(title)
    if cuda:
        import torch
        la = title_number_of_characters.reshape(-1).to(torch.uint8).contiguous()
        lb = truth_number_of_characters.reshape(-1).to(torch.uint8).contiguous()
        counts = truth_words_counts.reshape(n_pairs, NUMBER_OF_WORDS_FEATURES).contiguous()
        if str(counts.dtype) not in ('torch.uint32', 'torch.int32'):
            raise TypeError('truth_words_counts must be uint32 (or int32 bit pattern) on the device')
        title, title_truth = title.contiguous(), title_truth.contiguous()
        if response is None:
            response = torch.empty((n_pairs, FEATURES_COUNT), dtype=torch.float32, device=title.device)
    else:
        la = np.ascontiguousarray(np.asarray(title_number_of_characters).reshape(-1), dtype=np.uint8)
        lb = np.ascontiguousarray(np.asarray(truth_number_of_characters).reshape(-1), dtype=np.uint8)
        counts = np.ascontiguousarray(np.asarray(truth_words_counts).reshape(n_pairs, NUMBER_OF_WORDS_FEATURES),
                                      dtype=np.uint32)
        title, title_truth = np.ascontiguousarray(title), np.ascontiguousarray(title_truth)
        if response is None:
            response = np.empty((n_pairs, FEATURES_COUNT), dtype=np.float32)
    if la.shape[0] != n_pairs or lb.shape[0] != n_pairs:
        raise ValueError('length arrays must have one entry per pair')
    out = response if response.ndim == 2 else response.reshape(1, -1)
    if tuple(out.shape) != (n_pairs, FEATURES_COUNT) or 'float32' not in str(out.dtype):
        raise ValueError(f'response must be float32 [{n_pairs}, {FEATURES_COUNT}]')
    direct = out.is_contiguous() if cuda else out.flags['C_CONTIGUOUS']
    target = out if direct else (out.contiguous() if cuda else np.ascontiguousarray(out))
    nat.check(nat.lib.ds_construct_features(
        nat.ptr(la), nat.ptr(lb), nat.ptr(title), nat.ptr(title_truth), stride, nat.ptr(counts), int(space_code),
        int(number_of_truth_titles), n_pairs, nat.ptr(target), nat.stream_for(title, target)))
    if not direct:
        out[...] = target
    return response


def fast_levenshtein_ratio_batch(sequences, sequences_to_compare_against, lengths, lengths_to_compare_against,
                                 with_distance=False):
    """fast_levenshtein_ratio (feature_engineering.py:25-63) over P pairs in the padded [P, l] layout."""
    a = _as_2d_u8(sequences, 'sequences')
    b = _as_2d_u8(sequences_to_compare_against, 'sequences_to_compare_against')
    n, stride = a.shape
    cuda = _is_cuda(a)
    if cuda:
        import torch
        la, lb = lengths.to(torch.uint8).contiguous(), lengths_to_compare_against.to(torch.uint8).contiguous()
        out = torch.empty(n, dtype=torch.uint8, device=a.device)
        dist = torch.empty(n, dtype=torch.int16, device=a.device) if with_distance else None
        a, b = a.contiguous(), b.contiguous()
    else:
        la = np.ascontiguousarray(lengths, dtype=np.uint8)
        lb = np.ascontiguousarray(lengths_to_compare_against, dtype=np.uint8)
        out = np.empty(n, dtype=np.uint8)
        dist = np.empty(n, dtype=np.uint16) if with_distance else None
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    nat.check(nat.lib.ds_indel_ratio_u8(nat.ptr(a), nat.ptr(b), stride, nat.ptr(la), nat.ptr(lb), n, nat.ptr(out),
                                        nat.ptr(dist), nat.stream_for(a, out)))
    return (out, dist) if with_distance else out


def fast_levenshtein_ratio(sequence, sequence_to_compare_against):
    """Scalar signature of the reference (feature_engineering.py:25): two uint8 code arrays -> uint8."""
    a = np.ascontiguousarray(sequence, dtype=np.uint8)
    b = np.ascontiguousarray(sequence_to_compare_against, dtype=np.uint8)
    if a.shape[0] + b.shape[0] == 0:
        raise ZeroDivisionError('division by zero')     # what the reference's jitted function raises
    stride = max(a.shape[0], b.shape[0], 1)
    pa = np.zeros((1, stride), dtype=np.uint8)
    pb = np.zeros((1, stride), dtype=np.uint8)
    pa[0, :a.shape[0]] = a
    pb[0, :b.shape[0]] = b
    if stride > MAX_CHARACTERS_ALLOWED_IN_THE_TITLE:
        raise ValueError('sequences longer than 255 are outside the uint8 length domain of the batched kernel')
    out = fast_levenshtein_ratio_batch(pa, pb, np.array([a.shape[0]], np.uint8), np.array([b.shape[0]], np.uint8))
    return np.uint8(out[0])


def encode_title(title):
    """FeatureEngineering.encode_title (feature_engineering.py:298-307)."""
    out = np.zeros(MAX_CHARACTERS_ALLOWED_IN_THE_TITLE, dtype=np.uint8)
    codes = [ENCODING[ch] for ch in title[:MAX_CHARACTERS_ALLOWED_IN_THE_TITLE]]
    out[:len(codes)] = codes
    return out


def encode_titles(titles):
    """Compact table form used by the *_pairs kernels: (codes uint8[total], offsets int64[n+1])."""
    lengths = np.fromiter(map(len, titles), dtype=np.int64, count=len(titles))
    if lengths.size and lengths.max() > MAX_CHARACTERS_ALLOWED_IN_THE_TITLE:
        titles = [t[:MAX_CHARACTERS_ALLOWED_IN_THE_TITLE] for t in titles]
        np.minimum(lengths, MAX_CHARACTERS_ALLOWED_IN_THE_TITLE, out=lengths)
    joined = ''.join(titles)
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, code in ENCODING.items():
        lut[ord(ch)] = code
    raw = np.frombuffer(joined.encode('latin-1', 'replace'), dtype=np.uint8)
    codes = lut[raw]
    if codes.size and codes.max() == 255:
        bad = chr(int(raw[np.argmax(codes == 255)]))
        raise KeyError(bad)                              # encode_title would fail on self.encoding.get -> None
    offsets = np.zeros(len(titles) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    return codes, offsets


def get_truth_words_counts(title, words_counter):
    """FeatureEngineering.get_truth_words_counts (feature_engineering.py:309-319)."""
    out = np.zeros(NUMBER_OF_WORDS_FEATURES, dtype=np.uint32)
    counts = [words_counter.get(w) for w in title.split()][:NUMBER_OF_WORDS_FEATURES]
    out[:len(counts)] = counts
    return out


def construct_features_pairs(title_table, truth_table, truth_words_counts, title_index, truth_index, space_code,
                             number_of_truth_titles, response=None):
    """Same features from compact tables: `title_table` / `truth_table` = (codes, offsets) from
    `encode_titles`, `truth_words_counts` uint32 [n_truth_titles, 15], pair p = (title_index[p],
    truth_index[p]).  Avoids the [P,255] materialisation of predict.py:199-204."""
    bytes_a, off_a = title_table
    bytes_b, off_b = truth_table
    n_pairs = int(title_index.shape[0])
    cuda = _is_cuda(title_index)
    if response is None:
        if cuda:
            import torch
            response = torch.empty((n_pairs, FEATURES_COUNT), dtype=torch.float32, device=title_index.device)
        else:
            response = np.empty((n_pairs, FEATURES_COUNT), dtype=np.float32)
    nat.expect(title_index, 'int32', 'title_index')
    nat.expect(truth_index, 'int32', 'truth_index')
    nat.check(nat.lib.ds_construct_features_pairs(
        nat.ptr(bytes_a), nat.ptr(off_a), int(off_a.shape[0]) - 1, nat.ptr(bytes_b), nat.ptr(off_b), int(off_b.shape[0]) - 1,
        nat.ptr(truth_words_counts), nat.ptr(title_index), nat.ptr(truth_index), int(space_code),
        int(number_of_truth_titles), n_pairs, nat.ptr(response), nat.stream_for(title_index, response)))
    return response
