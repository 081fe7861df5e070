"""Deterministic synthetic titles with the example data set's length / word / trigram statistics
(SURVEY.md 8(d)).  Used by bench.py and the large parity tests; nothing here is on the timed path.

Truth titles: words per title from the example histogram.  The LAST word of a title comes from the example's
last-word table (the company forms `ltd` 27 %, `limited` 26 %, `bv` 16 % ... keep their share of TITLES, so the
commonest trigrams sit in 27 % of the titles like in the example), every other word from the table of the
remaining positions - the full 25,293-word table.  The table's tail (words seen <= 5 times in the example:
30.6 % of all word occurrences) is replaced by FRESH pseudo-words from an order-2 character model of those
tail words, so that hundreds of thousands of distinct titles come out while trigram document frequencies stay
Zipf-like; repeated titles get one more fresh word (the example's titles are 99.9 % distinct).
Measured against the example (30,000 truth / 10,000 test titles): top trigram in 27 % of the titles (example
27 %), postings touched per query 0.87 N (example 0.89 N), mean length 23.4 (23.4) - `workload_statistics`.
Test titles: `matched` of them are a truth title with 1-2 typing edits, the rest are fresh titles.
`tile_titles` grows a REAL title list (the example data) to any size instead: every copy beyond the first
carries 1-2 typing edits.  All titles are already in `transform_title` normal form (common.py:20-47).
"""
import os

import numpy as np

_STATS = None
TRUTH_SEED = 20240501
TEST_SEED = 20240502
PAIRS_SEED = 20240503
_CHUNK = 1 << 20


def _stats():
    global _STATS
    if _STATS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'example_word_stats.npz')
        raw = np.load(path)
        last = raw['count_last'].astype(np.float64)
        other = raw['count_other'].astype(np.float64)
        trans = raw['transitions'].astype(np.float64)
        n_codes = trans.shape[0]
        totals = trans.sum(axis=2, keepdims=True)
        # per state (a, b): cumulative distribution of the next character, laid out so that one searchsorted over
        # `state + u` (u uniform in [0, 1)) samples every word of a batch at once; unseen states end the word
        cdf = np.cumsum(np.where(totals > 0, trans / np.maximum(totals, 1.0), 0.0), axis=2)
        cdf[..., 0] = np.where(totals[..., 0] > 0, cdf[..., 0], 1.0)
        cdf = np.maximum.accumulate(cdf, axis=2)
        cdf[..., -1] = 1.0
        _STATS = dict(
            words=np.array([str(w) for w in raw['words']], dtype=object), last_p=last / last.sum(), other_p=other / other.sum(),
            is_tail=(last + other) <= float(raw['tail_count']),
            words_per_title=raw['words_per_title'].astype(np.float64) / raw['words_per_title'].sum(),
            word_lengths=raw['word_lengths'].astype(np.float64) / raw['word_lengths'].sum(),
            letters=np.array([''] + [str(c) for c in raw['letters']], dtype=object), n_codes=n_codes,
            next_cdf=(cdf + np.arange(n_codes * n_codes, dtype=np.float64).reshape(n_codes, n_codes, 1)).reshape(-1))
    return _STATS


def _pseudo_words(rng, n, max_length=24):
    """n fresh words from the order-2 character model of the example's tail words."""
    st = _stats()
    n_codes = st['n_codes']
    out = []
    for c0 in range(0, n, _CHUNK):
        m = min(_CHUNK, n - c0)
        a = np.zeros(m, dtype=np.int64)
        b = np.zeros(m, dtype=np.int64)
        codes = np.zeros((m, max_length), dtype=np.int8)
        length = np.zeros(m, dtype=np.int64)
        alive = np.arange(m)
        for _ in range(max_length):
            if alive.size == 0:
                break
            state = a[alive] * n_codes + b[alive]
            c = np.searchsorted(st['next_cdf'], state + rng.random(alive.size), side='right') - state * n_codes
            c = np.clip(c, 0, n_codes - 1)
            go = alive[c > 0]
            codes[go, length[go]] = c[c > 0]
            length[go] += 1
            a[go] = b[go]
            b[go] = c[c > 0]
            alive = go
        flat = ''.join(st['letters'][codes[codes > 0]].tolist())
        ends = np.cumsum(length)
        starts = ends - length
        out.extend(flat[s:e] if e > s else 'x' for s, e in zip(starts, ends))
    return out


def _normalise(title):
    title = ' '.join(title.split())[:255].strip()
    return title.rjust(3, '0') if len(title) < 3 else title


def generate_titles(n, seed, distinct=True):
    """n titles from the word model (see the module docstring)."""
    st = _stats()
    rng = np.random.default_rng(seed)
    out = []
    for c0 in range(0, n, _CHUNK):
        m = min(_CHUNK, n - c0)
        n_words = np.maximum(rng.choice(len(st['words_per_title']), size=m, p=st['words_per_title']), 1)
        total = int(n_words.sum())
        ends = np.cumsum(n_words)
        starts = ends - n_words
        picks = rng.choice(len(st['words']), size=total, p=st['other_p'])
        picks[ends - 1] = rng.choice(len(st['words']), size=m, p=st['last_p'])
        words = st['words'][picks]
        fresh = np.nonzero(st['is_tail'][picks])[0]
        if fresh.size:
            words[fresh] = np.array(_pseudo_words(rng, fresh.size), dtype=object)
        words = words.tolist()
        out.extend(_normalise(' '.join(words[s:e])) for s, e in zip(starts, ends))
    if distinct:
        seen = set()
        repeated = []
        for i, title in enumerate(out):
            if title in seen:
                repeated.append(i)
            else:
                seen.add(title)
        if repeated:
            extra = _pseudo_words(rng, len(repeated))
            for i, word in zip(repeated, extra):
                out[i] = _normalise(word + ' ' + out[i])
    return out


_KEYS = 'abcdefghijklmnopqrstuvwxyz0123456789'


def _misspell(rng_ints, rng_pos, title, k):
    """1-2 typing edits in the spirit of feature_engineering_prepare.py:60-173 (add / remove / replace /
    space / swap), driven by pre-drawn integers."""
    chars = list(title)
    for e in range(k):
        if len(chars) < 2:
            break
        op = rng_ints[e] % 5
        pos = rng_pos[e] % len(chars)
        key = _KEYS[(rng_ints[e] // 5) % len(_KEYS)]
        if op == 0:
            chars.insert(pos, key)
        elif op == 1:
            del chars[pos]
        elif op == 2:
            chars[pos] = key
        elif op == 3:
            chars.insert(pos, ' ')
        else:
            j = min(pos + 1, len(chars) - 1)
            chars[pos], chars[j] = chars[j], chars[pos]
    return _normalise(''.join(chars))


def generate_truth_titles(n, seed=TRUTH_SEED):
    return generate_titles(n, seed)


def generate_test_titles(truth_titles, n, seed=TEST_SEED, matched=0.6):
    rng = np.random.default_rng(seed)
    from_truth = rng.random(n) < matched
    source = rng.integers(0, len(truth_titles), size=n)
    n_edits = rng.integers(1, 3, size=n)
    ints = rng.integers(0, 1 << 30, size=(n, 2))
    pos = rng.integers(0, 1 << 30, size=(n, 2))
    fresh = generate_titles(int((~from_truth).sum()), seed + 1)
    out, f = [], 0
    for i in range(n):
        if from_truth[i]:
            out.append(_misspell(ints[i], pos[i], truth_titles[source[i]], int(n_edits[i])))
        else:
            out.append(fresh[f])
            f += 1
    return out, np.where(from_truth, source, -1)


def tile_titles(base_titles, n, seed):
    """Grows a real title list to n titles: copy i is base_titles[i % len]; every copy beyond the first pass carries
    1-2 typing edits in front of its last word (the company form - ltd, limited, bv - stays, like in distinct real
    companies), so the rows stay distinct while word and trigram statistics stay the data set's own."""
    rng = np.random.default_rng(seed)
    base = len(base_titles)
    n_edits = rng.integers(1, 3, size=n)
    ints = rng.integers(0, 1 << 30, size=(n, 2))
    pos = rng.integers(0, 1 << 30, size=(n, 2))
    out = []
    for i in range(n):
        title = base_titles[i % base]
        if i >= base:
            body, space, last = title.rpartition(' ')
            if body:
                title = _normalise(_misspell(ints[i], pos[i], body, int(n_edits[i])) + space + last)
            else:
                title = _misspell(ints[i], pos[i], title, int(n_edits[i]))
        out.append(title)
    return out


def workload_statistics(enc):
    """Density of an encoded workload (`encode.encode_canonical*` output, host or device arrays): the figures SURVEY.md
    8(d) pins the generator to.  postings_hit_per_query_over_n = mean over queries of sum(df of its trigrams) / N - the
    number of (row, column) incidences a posting-list scan touches per query (example data: 0.89)."""
    to_np = lambda x: x.cpu().numpy() if hasattr(x, 'cpu') else np.asarray(x)   # noqa: E731
    t_cols, q_cols, q_ptr, t_ptr = (to_np(enc[key]) for key in ('t_cols', 'q_cols', 'q_ptr', 't_ptr'))
    n_truth, n_queries = t_ptr.shape[0] - 1, q_ptr.shape[0] - 1
    df = np.bincount(t_cols, minlength=int(to_np(enc['idf64']).shape[0]))
    present = np.sort(df[df > 0])
    return dict(top_trigram_df_share=float(present[-1] / n_truth) if present.size else 0.0,
                median_trigram_df=float(np.median(present)) if present.size else 0.0,
                singleton_trigrams=int((present == 1).sum()), vocab=int(df.shape[0]),
                mean_trigrams_per_truth_title=float(t_cols.shape[0] / max(1, n_truth)),
                postings_hit_per_query_over_n=float(df[q_cols].sum() / max(1, n_queries) / max(1, n_truth)))


def generate_long_titles(n, seed, low=65, high=128):
    """Stress slice of BASELINE config 4: titles with lengths uniform in [low, high]."""
    rng = np.random.default_rng(seed)
    pool = generate_titles(n * 6, seed + 7)
    targets = rng.integers(low, high + 1, size=n)
    out, p = [], 0
    for t in targets:
        parts, length = [], 0
        while length < t:
            w = pool[p % len(pool)]
            p += 1
            parts.append(w)
            length += len(w) + 1
        out.append(_normalise(' '.join(parts)[:int(t)]))
    return out
