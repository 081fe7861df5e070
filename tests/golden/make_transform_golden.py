"""Mints tests/golden/transform_titles.npz from the reference's own `transform_title` (common.py:20-47).

Container only (imports /root/reference through oracle/ref_import.py):  python tests/golden/make_transform_golden.py
Inputs: the raw `title` column of the example data (first 3,000 truth + 1,500 test rows), the reference's own test
vector (doppelspeller/tests/test_common.py:17) and seeded nasties: accents, ligatures, dashes, every ASCII
white-space character, CJK / emoji, titles that shrink below 3 characters, titles beyond 255 characters with white
space around the cut.  Stored as UTF-8 blobs + offsets (inputs) and latin-1 blobs + offsets (outputs).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def nasty_titles(rng, count):
    pools = [
        'abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789',
        ' \t\n\r\x0b\x0c\x1c\x1d\x1e\x1f  --__..,,;;&&**()[]{}\'"/\\|!?@#$%^+=~`<>',
        'àáâãäåçèéêëìíîïñòóôõöùúûüýÿÀÉÎÕÜŠšŽžŒœßøØđĐłŁæÆ',
        '≠≮≯KÅ;`́̈  ​—–­',
        '株式会社有限公司한국어ﬁﬂ①②Ⅷ😀🏢',
    ]
    out = []
    for i in range(count):
        length = int(rng.choice([0, 1, 2, 3, 5, 12, 30, 60, 120, 254, 255, 256, 257, 300, 700]))
        weights = rng.dirichlet(np.ones(len(pools)) * 0.6)
        chars = []
        for _ in range(length):
            pool = pools[int(rng.choice(len(pools), p=weights))]
            chars.append(pool[int(rng.integers(len(pool)))])
        out.append(''.join(chars))
    # white space exactly around the 255 cut and tiny results
    out += ['a' * 254 + ' ' + 'b' * 10, 'a' * 253 + '  \t ' + 'b' * 10, ' ' * 300 + 'ab', '\t\tx\n', '--', '', ' ', 'é', 'Ab', 'a-b',
            'a\tb  c\n\nd', '  multiple   spaces\t\ttabs  ', 'x' * 255, 'x' * 256, ('ab ' * 100)]
    return out


def main():
    from oracle import ref_import
    common = ref_import.import_reference().common
    import pandas as pd
    data_dir = ref_import.stage_example_data()
    titles = []
    for name, rows in (('example_truth.csv', 3000), ('example_test.csv', 1500)):
        frame = pd.read_csv(os.path.join(data_dir, name), delimiter='|', nrows=rows)
        titles += [str(x) for x in frame['name'].tolist()]
    titles.append('''LKJblksd skjasl dfkjf &* 8*&&&8 GGdjsdkj--sdsd-"sdi..//' d'  k   bkjh77_asda33''')
    titles += nasty_titles(np.random.default_rng(20240504), 2500)
    import logging
    logging.disable(logging.CRITICAL)
    outputs = [common.transform_title(t) for t in titles]

    def blob(strings, encoding):
        encoded = [s.encode(encoding, 'surrogatepass') if encoding.startswith('utf') else s.encode(encoding) for s in strings]
        offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
        np.cumsum([len(e) for e in encoded], out=offsets[1:])
        return np.frombuffer(b''.join(encoded), dtype=np.uint8).copy(), offsets
    in_bytes, in_off = blob(titles, 'utf-8')
    out_bytes, out_off = blob(outputs, 'latin-1')
    target = os.path.join(ROOT, 'tests', 'golden', 'transform_titles.npz')
    np.savez_compressed(target, in_bytes=in_bytes, in_off=in_off, out_bytes=out_bytes, out_off=out_off)
    print(f'{len(titles)} titles -> {target} ({os.path.getsize(target)} bytes)')


if __name__ == '__main__':
    main()
