"""Deterministic synthetic titles with the example data set's length / word / trigram statistics
(SURVEY.md 8(d)).  Used by bench.py and the large parity tests; nothing here is on the timed path.

Truth titles: words per title from the example histogram; each word is drawn from the example
word-frequency table (heavy head: ltd, limited, bv ...) or, with probability `fresh`, is a fresh
pseudo-word (length from the example word-length distribution, letters from its unigram table) so that
hundreds of thousands of distinct titles with a Zipf-like trigram document frequency come out.
Test titles: `matched` of them are a truth title with 1-2 typing edits, the rest are fresh titles.
All titles are already in `transform_title` normal form (common.py:20-47).
"""
import os

import numpy as np

_STATS = None
TRUTH_SEED = 20240501
TEST_SEED = 20240502
PAIRS_SEED = 20240503


def _stats():
    global _STATS
    if _STATS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'example_word_stats.npz')
        raw = np.load(path)
        words = [str(w) for w in raw['words']]
        counts = raw['word_counts'].astype(np.float64)
        tail = float(raw['total_word_occurrences']) - counts.sum()       # mass of the words beyond the table
        _STATS = dict(
            words=np.array(words, dtype=object), word_p=counts / counts.sum(), tail_fraction=tail / (tail + counts.sum()),
            words_per_title=raw['words_per_title'].astype(np.float64) / raw['words_per_title'].sum(),
            word_lengths=raw['word_lengths'].astype(np.float64) / raw['word_lengths'].sum(),
            letters=np.array([str(c) for c in raw['letters']], dtype=object),
            letter_p=raw['letter_counts'].astype(np.float64) / raw['letter_counts'].sum())
    return _STATS


def _pseudo_words(rng, n):
    st = _stats()
    lengths = np.maximum(rng.choice(len(st['word_lengths']), size=n, p=st['word_lengths']), 1)
    letters = rng.choice(st['letters'], size=int(lengths.sum()), p=st['letter_p'])
    flat = ''.join(letters)
    ends = np.cumsum(lengths)
    starts = ends - lengths
    return [flat[s:e] for s, e in zip(starts, ends)]


def _normalise(title):
    title = ' '.join(title.split())[:255].strip()
    return title.rjust(3, '0') if len(title) < 3 else title


def generate_titles(n, seed, fresh=0.3):
    """n titles from the word model (fresh = share of pseudo-words on top of the table's own tail share)."""
    st = _stats()
    rng = np.random.default_rng(seed)
    n_words = np.maximum(rng.choice(len(st['words_per_title']), size=n, p=st['words_per_title']), 1)
    total = int(n_words.sum())
    is_fresh = rng.random(total) < (fresh + (1.0 - fresh) * st['tail_fraction'])
    picks = rng.choice(len(st['words']), size=total, p=st['word_p'])
    words = st['words'][picks]
    n_fresh = int(is_fresh.sum())
    if n_fresh:
        words[np.nonzero(is_fresh)[0]] = np.array(_pseudo_words(rng, n_fresh), dtype=object)
    ends = np.cumsum(n_words)
    starts = ends - n_words
    words = words.tolist()
    return [_normalise(' '.join(words[s:e])) for s, e in zip(starts, ends)]


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
