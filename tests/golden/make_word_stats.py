"""Derives the small statistics table the synthetic title generator bootstraps from
(SURVEY.md 8(d)): the full word-frequency table of the example truth titles (split by position: last word of a
title / any other position, so that the company-form suffixes `ltd`, `limited`, `bv` ... keep their share of
TITLES), the words-per-title histogram, the word-length and letter distributions and an order-2 character model
of the table's tail (the words the generator replaces by fresh pseudo-words).  Container only (reads the golden
example_titles.npz, itself minted from the reference's example data set by make_golden.py).

    python tests/golden/make_word_stats.py   ->  doppelspeller_b200/data/example_word_stats.npz
"""
import os
from collections import Counter

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TAIL_COUNT = 5          # words seen at most this often (30.6 % of all word occurrences) are "tail" words
ALPHABET = 'abcdefghijklmnopqrstuvwxyz0123456789'


def main():
    titles = [str(t) for t in np.load(os.path.join(HERE, 'example_titles.npz'))['truth_titles']]
    words_per_title = Counter()
    count_last, count_other = Counter(), Counter()
    letters = Counter()
    lengths = Counter()
    for title in titles:
        words = title.split()
        words_per_title[min(len(words), 12)] += 1
        count_last[words[-1]] += 1
        count_other.update(words[:-1])
        for w in words:
            lengths[min(len(w), 24)] += 1
            letters.update(w)
    total = count_last + count_other
    vocabulary = [w for w, _ in total.most_common()]
    # order-2 character model of the tail words: transitions[a, b, c] = occurrences of c after (a, b); 0 = word boundary
    code = {ch: i + 1 for i, ch in enumerate(ALPHABET)}
    transitions = np.zeros((len(ALPHABET) + 1,) * 3, dtype=np.uint32)
    for w in vocabulary:
        if total[w] > TAIL_COUNT:
            continue
        s = [0, 0] + [code[ch] for ch in w] + [0]
        for a, b, c in zip(s, s[1:], s[2:]):
            transitions[a, b, c] += total[w]
    np.savez_compressed(
        os.path.join(ROOT, 'doppelspeller_b200', 'data', 'example_word_stats.npz'),
        words=np.array(vocabulary), count_last=np.array([count_last.get(w, 0) for w in vocabulary], dtype=np.int64),
        count_other=np.array([count_other.get(w, 0) for w in vocabulary], dtype=np.int64), tail_count=np.int64(TAIL_COUNT),
        words_per_title=np.array([words_per_title.get(i, 0) for i in range(13)], dtype=np.int64),
        word_lengths=np.array([lengths.get(i, 0) for i in range(25)], dtype=np.int64),
        letters=np.array(list(ALPHABET)), letter_counts=np.array([letters.get(ch, 0) for ch in ALPHABET], dtype=np.int64),
        transitions=transitions)
    tail = sum(c for c in total.values() if c <= TAIL_COUNT)
    print(len(total), sum(total.values()), f'tail share {tail / sum(total.values()):.3f}', total.most_common(6))


if __name__ == '__main__':
    main()
