"""Derives the small statistics table the synthetic title generator bootstraps from
(SURVEY.md 8(d)): word frequencies, words-per-title histogram, word-length and letter distributions of
the example truth titles.  Container only (reads the golden example_titles.npz, itself minted from the
reference's example data set by make_golden.py).

    python tests/golden/make_word_stats.py   ->  doppelspeller_b200/data/example_word_stats.npz
"""
import os
from collections import Counter

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TOP_WORDS = 6000


def main():
    titles = np.load(os.path.join(HERE, 'example_titles.npz'))['truth_titles']
    words_per_title = Counter()
    word_freq = Counter()
    letters = Counter()
    lengths = Counter()
    for title in titles:
        words = str(title).split()
        words_per_title[min(len(words), 12)] += 1
        word_freq.update(words)
        for w in words:
            lengths[min(len(w), 24)] += 1
            letters.update(w)
    top = word_freq.most_common(TOP_WORDS)
    alphabet = 'abcdefghijklmnopqrstuvwxyz0123456789'
    np.savez_compressed(
        os.path.join(ROOT, 'doppelspeller_b200', 'data', 'example_word_stats.npz'),
        words=np.array([w for w, _ in top]), word_counts=np.array([c for _, c in top], dtype=np.int64),
        total_word_occurrences=np.int64(sum(word_freq.values())), distinct_words=np.int64(len(word_freq)),
        words_per_title=np.array([words_per_title.get(i, 0) for i in range(13)], dtype=np.int64),
        word_lengths=np.array([lengths.get(i, 0) for i in range(25)], dtype=np.int64),
        letters=np.array(list(alphabet)), letter_counts=np.array([letters.get(ch, 0) for ch in alphabet], dtype=np.int64))
    print(len(word_freq), sum(word_freq.values()), top[:8], dict(words_per_title))


if __name__ == '__main__':
    main()
